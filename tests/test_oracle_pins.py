"""CPU: the oracle against every known-answer value that survives in the reference (tests/golden/pins.json,
SURVEY.md 8c) and against analytic identities.  No GPU."""
import json
import os

import numpy as np
import pytest

from conftest import relerr

HERE = os.path.dirname(os.path.abspath(__file__))
PINS = json.load(open(os.path.join(HERE, "golden", "pins.json")))


def test_le_uniaxial_pin(oracle):
    p = PINS["le_uniaxial"]
    s, D = oracle.le_stress(p["E"], p["nu"], p["eps"])
    assert abs(s[0] / 1e6 - p["sigma11_MPa"]) < 1e-6
    assert abs(s[1] / 1e6 - p["sigma22_MPa"]) < 1e-6 and abs(s[2] / 1e6 - p["sigma22_MPa"]) < 1e-6


def test_le_pure_shear_pin(oracle):
    p = PINS["le_pure_shear"]
    s, _ = oracle.le_stress(p["E"], p["nu"], [0, 0, 0, p["gamma12"] / 2, 0, 0])
    assert abs(s[3] / 1e6 - p["sigma12_MPa"]) < 5e-3


def test_le_tangent_identity(oracle):
    eps = np.array(PINS["le_tangent_identity"]["eps"])
    s, D = oracle.le_stress(200e9, 0.3, eps)
    voigt = eps.copy(); voigt[3:] *= 2.0          # engineering shears
    assert relerr(D @ voigt, s) < 1e-15


def test_pp_uniaxial_pin(oracle):
    p = PINS["pp_uniaxial"]
    s, D, st, plastic = oracle.pp_stress([p["E"], p["nu"], p["sigma_y"], p["H"]], [p["eps11"], 0, 0, 0, 0, 0])
    assert plastic
    assert abs(s[0] / 1e6 - p["sigma11_MPa"]) < 1e-5
    assert abs(st[0] - p["eps_p11"]) < 1e-8
    assert abs(st[6] / 1e6 - p["alpha11_MPa"]) < 1e-5
    assert abs(st[12] - p["kappa"]) < 1e-8


def test_pp_on_yield_surface(oracle):
    par = [200e9, 0.3, 250e6, 1e9]
    rng = np.random.default_rng(0)
    for _ in range(20):
        eps = 4e-3 * rng.standard_normal(6)
        s, D, st, plastic = oracle.pp_stress(par, eps)
        if not plastic:
            continue
        r = s - st[6:12]
        r[:3] -= r[:3].mean()
        q = np.sqrt(1.5 * (np.sum(r[:3] ** 2) + 2 * np.sum(r[3:] ** 2)))
        assert abs(q / 1e6 - PINS["pp_on_surface"]["sigma_y_MPa"]) < 1e-6


def test_pp_elastic_step_keeps_state(oracle):
    par = [200e9, 0.3, 250e6, 1e9]
    s, D, st, plastic = oracle.pp_stress(par, [1e-4, 0, 0, 0, 0, 0])
    assert not plastic and np.all(st == 0)
    s2, _ = oracle.le_stress(200e9, 0.3, [1e-4, 0, 0, 0, 0, 0])
    assert relerr(s, s2) < 1e-15


def test_quadrature_pins(oracle):
    q = PINS["quadrature"]
    w, xi = oracle.quadrature(10)
    assert len(w) == 4 and np.all(w == q["gltet4_weight"]) and abs(w.sum() - q["tet_volume"]) < 1e-16
    a, b = (5 + 3 * np.sqrt(5.0)) / 20, (5 - np.sqrt(5.0)) / 20
    assert np.array_equal(xi, np.array([[a, b, b], [b, a, b], [b, b, a], [b, b, b]]))
    w, xi = oracle.quadrature(4)
    assert len(w) == 1 and w[0] == q["gltet1_weight"] and np.all(xi == 0.25)
    w, xi = oracle.quadrature(8)
    assert len(w) == 8 and w.sum() == q["hex_volume"] and np.all(np.abs(xi) == q["glhex8_point"])
    # first index fastest (src/quadrature/glquad.jl:12-15)
    assert xi[0].tolist() == [-q["glhex8_point"]] * 3 and xi[1, 0] > 0 and xi[1, 1] < 0 and xi[2, 1] > 0 and xi[4, 2] > 0


def _gltet15():
    # degree-5 rule (src/quadrature/gltet.jl GLTET15), used only to integrate N_i N_j (degree 4) for the mass pin
    s15 = np.sqrt(15.0)
    a = 0.25
    b1, b2 = (7 + s15) / 34, (7 - s15) / 34
    c1, c2 = (13 - 3 * s15) / 34, (13 + 3 * s15) / 34
    d, f = (5 - s15) / 20, (5 + s15) / 20
    w1, w2, w3, w4 = 8 / 405, (2665 - 14 * s15) / 226800, (2665 + 14 * s15) / 226800, 5 / 567
    pts = [(a, a, a), (b1, b1, b1), (b1, b1, c1), (b1, c1, b1), (c1, b1, b1), (b2, b2, b2), (b2, b2, c2), (b2, c2, b2), (c2, b2, b2),
           (d, d, f), (d, f, d), (f, d, d), (d, f, f), (f, d, f), (f, f, d)]
    ws = [w1] + [w2] * 4 + [w3] * 4 + [w4] * 6
    return np.array(ws), np.array(pts)


def test_tet10_mass_matrix_pin(oracle):
    """Pins the Tet10 shape functions AND their node ordering against the integer table of
    src/assembly/assembly.jl:139-149 (M = detJ/2520 * table for a constant metric)."""
    table = np.array(PINS["tet10_mass_times_2520"]["table"], dtype=float)
    w, pts = _gltet15()
    M = np.zeros((10, 10))
    for wi, p in zip(w, pts):
        N = oracle.shape_N(10, p)
        M += wi * np.outer(N, N)
    assert np.abs(M * 2520 - table).max() < 1e-11


@pytest.mark.parametrize("et", [4, 8, 10])
def test_partition_of_unity_and_derivative_consistency(oracle, et):
    rng = np.random.default_rng(1)
    for _ in range(5):
        xi = rng.random(3) * (0.3 if et != 8 else 1.0)
        N, dN = oracle.shape_N(et, xi), oracle.shape_dN(et, xi)
        assert abs(N.sum() - 1) < 1e-14 and np.abs(dN.sum(axis=0)).max() < 1e-14
        h = 1e-6
        for a in range(3):
            e = np.zeros(3); e[a] = h
            fd = (oracle.shape_N(et, xi + e) - oracle.shape_N(et, xi - e)) / (2 * h)
            assert np.abs(fd - dN[:, a]).max() < 1e-8


def test_nh_closed_form_matches_energy_derivatives(oracle):
    """S = 2 dpsi/dC and DD = 4 d2psi/dC2 (src/materials/neo_hookean.jl:222) by central differences of strain_energy."""
    la, mu = oracle.lame(3e6, 0.45)
    rng = np.random.default_rng(12345)
    F = np.eye(3) + 0.15 * rng.standard_normal((3, 3))
    Cm = F.T @ F
    Em = 0.5 * (Cm - np.eye(3))
    idx = [(0, 0), (1, 1), (2, 2), (0, 1), (1, 2), (0, 2)]
    Ev = np.array([Em[i, j] for i, j in idx])
    S, D = oracle.nh_stress(mu, la, Ev)

    def psi(Cmat):
        return oracle.nh_energy(mu, la, np.array([Cmat[i, j] for i, j in idx]))

    h = 1e-6
    for q, (i, j) in enumerate(idx):
        dC = np.zeros((3, 3)); dC[i, j] += 0.5; dC[j, i] += 0.5
        dpsi = (psi(Cm + h * dC) - psi(Cm - h * dC)) / (2 * h)
        assert abs(2 * dpsi - S[q]) < 1e-6 * np.abs(S).max()
        # tangent: dS = DD : dE, with dE = dC/2
        Sp, _ = oracle.nh_stress(mu, la, Ev + h * np.array([0.5 * (dC[a, b] + dC[b, a]) / 1 for a, b in idx]) * 0.5 * 2 / 2)
        Sm, _ = oracle.nh_stress(mu, la, Ev - h * np.array([0.5 * (dC[a, b] + dC[b, a]) / 1 for a, b in idx]) * 0.5 * 2 / 2)
        dE = 0.5 * np.array([0.5 * (dC[a, b] + dC[b, a]) for a, b in idx])       # tensor components of dE = dC/2
        voigt = dE.copy(); voigt[3:] *= 2.0
        assert np.abs((Sp - Sm) / (2 * h) - D @ voigt).max() < 1e-5 * np.abs(D).max() * np.abs(voigt).max()
    # S(E=0) = 0 and small-strain limit = Hooke (docs/book/neo_hookean_implementation.md:249-259)
    S0, D0 = oracle.nh_stress(mu, la, np.zeros(6))
    assert np.abs(S0).max() == 0.0
    _, Dle = oracle.le_stress(3e6, 0.45, np.zeros(6))
    assert relerr(D0, Dle) < 1e-14
    with pytest.raises(ValueError):
        oracle.nh_stress(mu, la, np.array([-0.6, 0, 0, 0, 0, 0]))   # C11 = -0.2 -> det <= 0


def _random_curved_tet10(rng):
    ref = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1], [.5, 0, 0], [.5, .5, 0], [0, .5, 0], [0, 0, .5], [.5, 0, .5], [0, .5, .5]], float)
    A = np.eye(3) + 0.2 * rng.standard_normal((3, 3))
    return ref @ A.T + 0.02 * rng.standard_normal((10, 3))


def test_element_stiffness_identities(oracle):
    """Ke symmetric, 6 rigid-body modes, Voigt form == block form (src/physics/assembly_helpers.jl:218)."""
    rng = np.random.default_rng(12345)
    X = _random_curved_tet10(rng)
    Km, Kg, f, _ = oracle.element(10, X)
    assert relerr(Km, Km.T) < 1e-14
    ev = np.linalg.eigvalsh(Km)
    assert np.sum(np.abs(ev) < 1e-12 * ev[-1]) == 6 and ev[6] > 1e-4 * ev[-1]
    assert relerr(oracle.element_block_form(10, X, 210e9, 0.3), Km) < 1e-14
    # infinitesimal rigid rotation
    W = np.array([[0, -1, 2], [1, 0, -3], [-2, 3, 0]], float) * 1e-3
    u = (X @ W.T).ravel()
    assert np.abs(Km @ u).max() < 1e-12 * np.abs(Km).max() * np.abs(u).max()
    # f_int = Km u for small strain LE
    _, _, f, _ = oracle.element(10, X, u=rng.standard_normal((10, 3)) * 1e-3)


def test_patch_test_linear_field(oracle, jf):
    """Linear displacement => constant strain => nodal forces vanish at interior nodes."""
    m = jf.mesh.tet10_kuhn(2, 2, 2)
    G = np.array([[1e-3, 2e-4, 0], [0, -5e-4, 3e-4], [1e-4, 0, 7e-4]])
    u = (m.coords @ G.T).ravel()
    y = oracle.matfree(10, m.coords, m.conn, u).reshape(-1, 3)
    interior = np.all((m.coords > 1e-9) & (m.coords < 1 - 1e-9), axis=1)
    assert interior.sum() > 0
    assert np.abs(y[interior]).max() < 1e-10 * np.abs(y).max()
    assert np.abs(y.sum(axis=0)).max() < 1e-9 * np.abs(y).max()


def test_csr_pattern_is_union_of_element_dof_products(oracle, jf):
    m = jf.mesh.tet10_kuhn(2, 1, 1)
    rowptr, colind = oracle.csr_pattern(10, m.n_nodes, m.conn)
    ref = [set() for _ in range(m.n_dofs)]
    for e in range(m.n_elems):
        g = [3 * (n - 1) + c for n in m.conn[e] for c in range(3)]     # src/assembly/problems.jl:476 (0-based)
        for r in g:
            ref[r].update(g)
    for r in range(m.n_dofs):
        assert colind[rowptr[r]:rowptr[r + 1]].tolist() == sorted(ref[r])


def test_assembled_equals_matrix_free_and_scipy(oracle, jf):
    import scipy.sparse as sp
    m = jf.mesh.tet10_kuhn(3, 2, 2)
    rowptr, colind, vals, f = oracle.assemble_csr(10, m.coords, m.conn)
    u = jf.mesh.test_vector(m.n_dofs)
    assert relerr(oracle.spmv(rowptr, colind, vals, u), oracle.matfree(10, m.coords, m.conn, u)) < 1e-13
    # independent COO -> CSR through scipy (stand-in for Julia's sparse(I,J,V))
    I, J, V = [], [], []
    for e in range(m.n_elems):
        Ke, _, _, _ = oracle.element(10, m.coords[m.conn[e] - 1])
        g = np.array([3 * (n - 1) + c for n in m.conn[e] for c in range(3)])
        I.append(np.repeat(g, 30)); J.append(np.tile(g, 30)); V.append(Ke.ravel())
    K = sp.coo_matrix((np.concatenate(V), (np.concatenate(I), np.concatenate(J))), shape=(m.n_dofs, m.n_dofs)).tocsr()
    K.sort_indices()
    assert np.array_equal(K.indptr, rowptr) and np.array_equal(K.indices, colind)
    assert relerr(K.data, vals) < 1e-13


def test_cg_solves_cantilever(oracle, jf):
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    m = jf.mesh.tet10_kuhn(4, 1, 1, 4.0, 1.0, 1.0)
    fixed = jf.mesh.clamp_dofs(m)
    rowptr, colind, vals, _ = oracle.assemble_csr(10, m.coords, m.conn, par=(210e9, 0.3))
    b = np.zeros(m.n_dofs); b[2::3] = -1e3
    x, it, res = oracle.cg_csr(rowptr, colind, vals, b, fixed_dofs=fixed, tol=1e-10, relative=True, max_iter=5000)
    assert it < 5000
    K = sp.csr_matrix((vals, colind, rowptr))
    free = np.setdiff1d(np.arange(m.n_dofs), fixed - 1)
    xd = np.zeros(m.n_dofs)
    xd[free] = spla.spsolve(K[free][:, free].tocsc(), b[free])     # elimination solve (src/solvers.jl:205-210)
    assert relerr(x, xd) < 1e-6
    assert np.all(x[fixed - 1] == 0)


def test_colouring_is_valid(oracle, jf):
    m = jf.mesh.tet10_kuhn(3, 3, 3)
    col, n = oracle.colouring(10, m.n_nodes, m.conn)
    assert n >= 24
    for c in range(n):
        nodes = m.conn[col == c].ravel()
        assert nodes.size == np.unique(nodes).size


def _linear_static_case():
    """Mesh, fixed dofs (1-based) and pin of examples/linear_static.jl:23-100,133 (fixture: tests/golden/make_fixtures.py)."""
    from juliafem.jl_b200 import mesh
    z = np.load(os.path.join(HERE, "golden", "linear_static_smp18.npz"))
    m = mesh.Mesh(10, z["coords"], z["conn"])
    nodes = np.union1d(mesh.nodes_at_plane(m, 1, 50.0, 6.0), mesh.find_nearest_nodes(m, [165.0, 88.0, 10.0], 3))
    fixed = np.sort((3 * (nodes[:, None] - 1) + np.arange(1, 4)[None, :]).ravel())
    return m, fixed, PINS["linear_static"]


def test_linear_static_end_to_end_pin(oracle):
    """The one reference-held number for the whole assemble -> eliminate -> solve path: the oracle's K (Tet10 GLTET4,
    3(n-1)+c numbering, (K+K')/2), consistent body load and the elimination of src/solvers.jl:205-210 with a direct
    solve reproduce examples/linear_static.jl:133 far inside the reference's own isapprox tolerance."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    m, fixed, pin = _linear_static_case()
    assert (m.n_nodes, m.n_elems) == (pin["n_nodes"], pin["n_tet10"]) and fixed.size == 3 * 1939
    rp, ci, vals, _ = oracle.assemble_csr(10, m.coords, m.conn, par=(pin["E"], pin["nu"]), symmetrise=True)
    K = sp.csr_matrix((vals, ci, rp))
    f = oracle.body_load(10, m.coords, m.conn, (pin["displacement_load_1"], 0.0, 0.0))
    free = np.setdiff1d(np.arange(m.n_dofs), fixed - 1)
    u = np.zeros(m.n_dofs)
    u[free] = spla.splu(K[free][:, free].tocsc()).solve(f[free])
    umax = np.linalg.norm(u.reshape(-1, 3), axis=1).max()
    assert abs(umax / pin["max_u_norm"] - 1.0) < 1e-9, umax
    # the literal (assigning) form of src/problems_elasticity.jl:418-425 does NOT give the reference's answer
    w, xi = oracle.quadrature(10)
    X = m.coords[m.conn - 1]
    J = np.einsum("ia,eib->eab", oracle.shape_dN(10, xi[-1]), X)
    flit = np.zeros(m.n_dofs)
    np.add.at(flit, 3 * (m.conn - 1), (w[-1] * np.linalg.det(J))[:, None] * oracle.shape_N(10, xi[-1])[None, :])
    ul = np.zeros(m.n_dofs)
    ul[free] = spla.splu(K[free][:, free].tocsc()).solve(flit[free])
    assert abs(np.linalg.norm(ul.reshape(-1, 3), axis=1).max() / pin["max_u_norm"] - 1.0) > 0.5


def test_body_load_total_and_python_mirror(oracle):
    from juliafem.jl_b200 import mesh
    m = mesh.tet10_kuhn(3, 2, 2, 3.0, 1.0, 2.0)
    f = oracle.body_load(10, m.coords, m.conn, (1.0, -2.0, 0.5))
    assert np.allclose(f.reshape(-1, 3).sum(0), np.array([1.0, -2.0, 0.5]) * 6.0, rtol=1e-13)
    h = mesh.hex8_lattice(4, 3, 3, 0.5)
    f = oracle.body_load(8, h.coords, h.conn, (0.0, 0.0, 9.81))
    assert abs(f[2::3].sum() - 9.81 * 1.5 * 1.0 * 1.0) < 1e-12


@pytest.mark.skipif(not os.path.exists("/root/reference/examples/linear_static/JuliaFEMSMP18.med"), reason="reference tree not present")
def test_med_reader_matches_committed_fixture():
    """h5lite + mesh.read_med on the reference's .med file == the committed npz (so the fixture is what the file holds)."""
    from juliafem.jl_b200 import mesh
    m = mesh.read_med("/root/reference/examples/linear_static/JuliaFEMSMP18.med")
    z = np.load(os.path.join(HERE, "golden", "linear_static_smp18.npz"))
    assert np.array_equal(m.coords, z["coords"]) and np.array_equal(m.conn, z["conn"])
    assert set(m.elem_sets) == {"OTHER"} and m.elem_sets["OTHER"].size == m.n_elems
    with pytest.raises(ValueError):
        mesh.read_med("/root/reference/examples/linear_static/JuliaFEMSMP18.med", mesh_name="nope")


def test_surface_load_identities(oracle):
    """Oracle restatement of src/problems_elasticity.jl:454-502: total traction force = t * area for Tri3 / Tri6 / Quad4,
    pressure acts against the normal dX/dxi1 x dX/dxi2, Tri3 (GLTRI1) equals the lumped area/3 rule of ext:368-416."""
    X = np.array([[0, 0, 0], [2, 0, 0], [0, 1.5, 0], [1, 0, 0], [1, 0.75, 0], [0, 0.75, 0], [2, 1.5, 0]], float)
    area_tri, t = 1.5, np.array([0.3, -1.0, 2.0])
    for ft, face, area in ((3, [1, 2, 3], area_tri), (6, [1, 2, 3, 4, 5, 6], area_tri), (4, [1, 2, 7, 3], 3.0)):
        f = oracle.surface_load(ft, X, [face], traction=t).reshape(-1, 3)
        assert np.allclose(f.sum(0), t * area, rtol=1e-14)
        p = oracle.surface_load(ft, X, [face], pressure=4.0).reshape(-1, 3)
        assert np.allclose(p.sum(0), [0, 0, -4.0 * area], rtol=1e-14)        # normal = +z, positive pressure pushes -z
    f3 = oracle.surface_load(3, X, [[1, 2, 3]], traction=t).reshape(-1, 3)
    assert np.allclose(f3[:3], np.outer(np.full(3, area_tri / 3), t), rtol=1e-14)
    # Tri6 consistent load of a flat straight-sided face: corners get 0, mid-side nodes a third each
    f6 = oracle.surface_load(6, X, [[1, 2, 3, 4, 5, 6]], traction=(0, 0, 1.0)).reshape(-1, 3)[:6, 2]
    assert np.allclose(f6, [0, 0, 0, 0.5, 0.5, 0.5], atol=1e-15)


def test_lsq_recovery_rule_and_exactness(oracle):
    """lsq_fit (src/problems_elasticity.jl:547-594).  With the element's default rule the Tet10 / Tet4 mass matrix is
    singular (documented deviation: the exact rule of the reference's own rule table is used instead); with that rule a
    linear displacement field gives its constant strain back at every node, and the stress is D : strain."""
    import scipy.sparse as sp
    from juliafem.jl_b200 import mesh
    z = np.load(os.path.join(HERE, "golden", "tet10_fixture.npz"))
    m = mesh.Mesh(10, z["coords"], z["conn"])
    c0 = (m.conn - 1).astype(np.int64)
    for rule, singular in ((oracle.quadrature(10), True), (oracle.quadrature_mass(10), False)):
        A = np.zeros((m.n_nodes, m.n_nodes))
        for wg, x in zip(*rule):
            N, dN = oracle.shape_N(10, x), oracle.shape_dN(10, x)
            J = np.einsum("ia,eib->eab", dN, m.coords[c0])
            np.add.at(A, (np.repeat(c0, 10, axis=1).ravel(), np.tile(c0, (1, 10)).ravel()),
                      ((wg * np.linalg.det(J))[:, None, None] * (N[:, None] * N[None, :])[None]).ravel())
        ev = np.linalg.eigvalsh(0.5 * (A + A.T))
        assert (ev[0] < 1e-12 * ev[-1]) == singular
    w, _ = oracle.quadrature_mass(10)
    assert abs(w.sum() - 1.0 / 6.0) < 1e-15
    G = np.array([[1e-3, 2e-4, 0], [0, -5e-4, 3e-4], [1e-4, 0, 2e-3]])
    u = (m.coords @ G.T).ravel()
    eps = 0.5 * (G + G.T)
    ev = np.array([eps[0, 0], eps[1, 1], eps[2, 2], eps[0, 1], eps[1, 2], eps[0, 2]])
    e = oracle.lsq_recover(10, m.coords, m.conn, u, "strain")
    assert np.abs(e - ev).max() < 1e-15
    s = oracle.lsq_recover(10, m.coords, m.conn, u, "stress", par=(200e9, 0.3))
    la, mu = oracle.lame(200e9, 0.3)
    sv = 2 * mu * ev + la * ev[:3].sum() * np.array([1, 1, 1, 0, 0, 0])
    assert relerr(s, np.tile(sv, (m.n_nodes, 1))) < 1e-12


class _HD:
    """Hyper-dual number a + b e1 + c e2 + d e1 e2 (e1^2 = e2^2 = 0): exact first and mixed second derivatives by forward-mode
    automatic differentiation -- what Tensors.hessian does with nested ForwardDiff duals (Tensors.jl 1.16.2, called at
    src/materials/neo_hookean.jl:222)."""
    __slots__ = ("a", "b", "c", "d")

    def __init__(self, a, b=0.0, c=0.0, d=0.0):
        self.a, self.b, self.c, self.d = a, b, c, d

    @staticmethod
    def lift(x):
        return x if isinstance(x, _HD) else _HD(float(x))

    def __add__(self, o):
        o = _HD.lift(o)
        return _HD(self.a + o.a, self.b + o.b, self.c + o.c, self.d + o.d)
    __radd__ = __add__

    def __neg__(self):
        return _HD(-self.a, -self.b, -self.c, -self.d)

    def __sub__(self, o):
        return self + (-_HD.lift(o))

    def __rsub__(self, o):
        return _HD.lift(o) + (-self)

    def __mul__(self, o):
        o = _HD.lift(o)
        return _HD(self.a * o.a, self.a * o.b + self.b * o.a, self.a * o.c + self.c * o.a,
                   self.a * o.d + self.b * o.c + self.c * o.b + self.d * o.a)
    __rmul__ = __mul__

    def fn(self, f0, f1, f2):
        """g(self) for a scalar function with value f0, first derivative f1, second derivative f2 at self.a"""
        return _HD(f0, f1 * self.b, f1 * self.c, f1 * self.d + f2 * self.b * self.c)

    def log(self):
        return self.fn(np.log(self.a), 1.0 / self.a, -1.0 / self.a ** 2)

    def sqrt(self):
        r = np.sqrt(self.a)
        return self.fn(r, 0.5 / r, -0.25 / (r * self.a))


def _nh_energy_ad(mu, la, C):
    """strain_energy(material, C) of src/materials/neo_hookean.jl:129-143, written on hyper-dual entries"""
    I1 = C[0][0] + C[1][1] + C[2][2]
    det = (C[0][0] * (C[1][1] * C[2][2] - C[1][2] * C[2][1]) - C[0][1] * (C[1][0] * C[2][2] - C[1][2] * C[2][0])
           + C[0][2] * (C[1][0] * C[2][1] - C[1][1] * C[2][0]))
    J = det.sqrt()
    lnJ = J.log()
    return (mu / 2.0) * (I1 - 3.0) - mu * lnJ + (la / 2.0) * (lnJ * lnJ)


def test_nh_closed_form_equals_forward_mode_ad_of_the_energy(oracle):
    """The reference obtains S = 2 dpsi/dC and DD = 4 d2psi/dC2 by automatic differentiation of strain_energy
    (Tensors.hessian(psi, C, :all), src/materials/neo_hookean.jl:205-231).  Restating exactly that -- forward-mode AD of the
    same energy expression, symmetric-tensor derivative convention (perturbations of C_ij and C_ji together, halved off the
    diagonal) -- must give the oracle's closed form to rounding, not just to finite-difference accuracy."""
    la, mu = oracle.lame(3e6, 0.45)
    rng = np.random.default_rng(2024)
    idx = [(0, 0), (1, 1), (2, 2), (0, 1), (1, 2), (0, 2)]
    for trial in range(5):
        F = np.eye(3) + 0.2 * rng.standard_normal((3, 3))
        Cm = F.T @ F
        Ev = np.array([0.5 * (Cm[i, j] - (i == j)) for i, j in idx])
        S, D = oracle.nh_stress(mu, la, Ev)

        def direction(i, j):          # basis of the symmetric tensors: dC = (e_i e_j' + e_j e_i') / 2
            dC = np.zeros((3, 3))
            dC[i, j] += 0.5
            dC[j, i] += 0.5
            return dC

        S_ad, D_ad = np.zeros(6), np.zeros((6, 6))
        for p, (i, j) in enumerate(idx):
            d1 = direction(i, j)
            for q, (k, l) in enumerate(idx):
                d2 = direction(k, l)
                C = [[_HD(Cm[a, b], d1[a, b], d2[a, b], 0.0) for b in range(3)] for a in range(3)]
                psi = _nh_energy_ad(mu, la, C)
                S_ad[p] = 2.0 * psi.b                     # S_ij = 2 dpsi/dC_ij
                D_ad[p, q] = 4.0 * psi.d                  # DD_ijkl = 4 d2psi/dC_ij dC_kl
        assert relerr(S, S_ad) < 1e-12
        assert relerr(D, D_ad) < 1e-12
        # the energy itself
        psi0 = _nh_energy_ad(mu, la, [[_HD(Cm[a, b]) for b in range(3)] for a in range(3)]).a
        assert abs(psi0 - oracle.nh_energy(mu, la, np.array([Cm[i, j] for i, j in idx]))) <= 1e-12 * abs(psi0)
