"""GPU: the steps either side of the solve (SURVEY.md 8 f2 / f4 and the partial rows a8b, a12, a16) through the C ABI, against
the oracle: consistent body / surface loads, reaction forces, least-squares stress recovery, the St. Venant-Kirchhoff
(finite_strain) path of the classic API and the CPU backend's penalty Dirichlet conditions."""
import numpy as np
import pytest

from conftest import curved_tet10, distorted_hex8, relerr

pytestmark = pytest.mark.gpu
LE = (210e9, 0.3)
TOL = 1e-12


@pytest.fixture(scope="module")
def L(jf):
    from juliafem.jl_b200 import _lib
    return _lib


def make(L, m, kind=0, par=LE, **opts):
    h = L.Handle(m.elem_type, m.coords, m.conn)
    for k, v in opts.items():
        h.set_option(k, v)
    h.set_material(kind, par)
    return h


@pytest.mark.parametrize("et", [10, 8, 4])
def test_body_load_matches_oracle(L, oracle, jf, et):
    m = {10: lambda: curved_tet10(jf.mesh, 4, 3, 2), 8: lambda: distorted_hex8(jf.mesh, 5), 4: lambda: jf.mesh.tet4_kuhn(4, 3, 2)}[et]()
    h = make(L, m)
    f = h.body_load((1.0, -2.5, 0.25))
    ref = oracle.body_load(et, m.coords, m.conn, (1.0, -2.5, 0.25))
    assert relerr(f, ref) < TOL
    # per-element loads, accumulated on top of an existing vector
    rng = np.random.default_rng(1)
    b = rng.standard_normal((m.n_elems, 3))
    f0 = rng.standard_normal(m.n_dofs)
    f2 = h.body_load(b, f=f0)
    assert relerr(f2 - f0, oracle.body_load(et, m.coords, m.conn, b)) < TOL
    assert np.array_equal(h.body_load(b, f=f0), f2)        # deterministic


def _boundary_faces(m, axis, value):
    """Faces of the volume mesh on the plane x_axis = value, in the reference's face node order (Tri6 for Tet10, Quad4 for Hex8)."""
    on = np.abs(m.coords[:, axis] - value) < 1e-12
    faces = []
    if m.elem_type == 10:
        tri = [(0, 1, 2, 4, 5, 6), (0, 1, 3, 4, 8, 7), (1, 2, 3, 5, 9, 8), (0, 2, 3, 6, 9, 7)]
        for c in m.conn:
            for t in tri:
                n = c[list(t)]
                if on[n - 1].all():
                    faces.append(n)
        return 6, np.array(faces)
    quad = [(0, 1, 2, 3), (4, 5, 6, 7), (0, 1, 5, 4), (1, 2, 6, 5), (2, 3, 7, 6), (3, 0, 4, 7)]
    for c in m.conn:
        for q in quad:
            n = c[list(q)]
            if on[n - 1].all():
                faces.append(n)
    return 4, np.array(faces)


def test_surface_traction_and_pressure_match_oracle(L, oracle, jf):
    for m in (jf.mesh.tet10_kuhn(4, 3, 2, 2.0, 1.5, 1.0), jf.mesh.hex8_lattice(5, 4, 3, 0.5)):
        top = m.coords[:, 2].max()
        ft, faces = _boundary_faces(m, 2, top)
        assert len(faces) > 0
        rng = np.random.default_rng(2)
        m.coords[:, 2] += 0.03 * np.sin(3 * m.coords[:, 0]) * (m.coords[:, 2] / top)      # curved top surface
        h = make(L, m)
        t = rng.standard_normal((len(faces), 3))
        p = rng.standard_normal(len(faces))
        f = h.surface_load(ft, faces, traction=t, pressure=p)
        ref = oracle.surface_load(ft, m.coords, faces, traction=t, pressure=p)
        assert relerr(f, ref) < TOL
        # uniform traction: total force = t * area; uniform pressure on a flat face acts against the normal
        hf = make(L, jf.mesh.hex8_lattice(5, 4, 3, 0.5))
        ft2, faces2 = _boundary_faces(jf.mesh.hex8_lattice(5, 4, 3, 0.5), 2, 1.0)
        fp = hf.surface_load(ft2, faces2, pressure=7.0)
        assert np.allclose(fp.reshape(-1, 3).sum(0), [0, 0, -7.0 * 2.0 * 1.5] if _normal_up(jf, faces2) else [0, 0, 7.0 * 2.0 * 1.5])
    # Tri3 consistent (GLTRI1) == the lumped area/3 rule of the reference's GPU kernel (ext/JuliaFEMCUDAExt.jl:368-416)
    m4 = jf.mesh.tet4_kuhn(3, 2, 2, 1.5)
    on = np.abs(m4.coords[:, 2] - m4.coords[:, 2].max()) < 1e-12
    tris = np.array([c[list(t)] for c in m4.conn for t in [(0, 1, 2), (0, 1, 3), (1, 2, 3), (0, 2, 3)] if on[c[list(t)] - 1].all()])
    f3 = make(L, m4).surface_load(3, tris, traction=(0.0, 0.0, -1e6))
    lump = np.zeros(m4.n_dofs)
    for tr in tris:
        X = m4.coords[tr - 1]
        a = 0.5 * np.linalg.norm(np.cross(X[1] - X[0], X[2] - X[0]))
        lump[3 * (tr - 1) + 2] += -1e6 * a / 3
    assert relerr(f3, lump) < TOL


def _normal_up(jf, faces):
    m = jf.mesh.hex8_lattice(5, 4, 3, 0.5)
    X = m.coords[faces[0] - 1]
    return np.cross(X[1] - X[0], X[3] - X[0])[2] > 0


def test_unsupported_surface_type_refused(L, jf):
    h = make(L, jf.mesh.tet4_kuhn(2, 2, 2))
    with pytest.raises(L.JfemError):
        h.surface_load(5, np.array([[1, 2, 3, 4, 5]]), traction=(0, 0, 1.0))


def test_reactions_balance_the_load(L, oracle, jf):
    m = jf.mesh.tet10_kuhn(6, 2, 2, 3.0, 1.0, 1.0)
    fixed = jf.mesh.clamp_dofs(m)
    h = make(L, m)
    h.set_dirichlet(fixed)
    b = h.body_load((0.0, 0.0, -7.8e4))
    x, it, res = h.cg(b, tol=1e-12, relative=True, max_iter=50000)
    la = h.reactions(x, b)
    assert np.all(la[np.setdiff1d(np.arange(m.n_dofs), fixed - 1)] == 0)
    # global equilibrium: reactions + applied load sum to zero, component by component
    tot = (la + b).reshape(-1, 3).sum(0)
    assert np.abs(tot).max() < 1e-8 * np.abs(b).sum()
    # and they are K u - f on the constrained rows (src/solvers.jl:211-216)
    Ku = oracle.matfree(10, m.coords, m.conn, x, par=LE)
    assert relerr(la[fixed - 1], (Ku - b)[fixed - 1]) < 1e-10


@pytest.mark.parametrize("et,field", [(10, "stress"), (10, "strain"), (8, "stress"), (4, "strain")])
def test_nodal_recovery_matches_oracle(L, oracle, jf, et, field):
    m = {10: lambda: curved_tet10(jf.mesh, 4, 3, 2), 8: lambda: distorted_hex8(jf.mesh, 5), 4: lambda: jf.mesh.tet4_kuhn(4, 3, 2)}[et]()
    u = 1e-3 * np.sin(2.0 * m.coords @ np.array([[1.0, 0.3, 0.0], [0.2, 1.0, 0.5], [0.0, 0.4, 1.0]])).ravel()
    h = make(L, m)
    x = h.nodal_recover(u, L.FIELD_STRESS if field == "stress" else L.FIELD_STRAIN)
    ref = oracle.lsq_recover(et, m.coords, m.conn, u, field, par=LE)
    assert relerr(x, ref) < 1e-9          # the reference factorises M; here Jacobi-PCG to 1e-14
    # a linear displacement field has constant strain: the fit reproduces it at every node
    G = np.array([[1e-3, 2e-4, 0], [0, -5e-4, 3e-4], [1e-4, 0, 2e-3]])
    e = h.nodal_recover((m.coords @ G.T).ravel(), L.FIELD_STRAIN)
    eps = 0.5 * (G + G.T)
    assert np.abs(e - np.array([eps[0, 0], eps[1, 1], eps[2, 2], eps[0, 1], eps[1, 2], eps[0, 2]])).max() < 1e-13


@pytest.mark.parametrize("et", [10, 8])
def test_stvk_finite_strain_path(L, oracle, jf, et):
    """props.finite_strain = true of the classic path (src/problems_elasticity.jl:255-332): Hooke on the Green-Lagrange
    strain.  f_int, Km (geometric_stiffness off) and Km + Kg (on) against the oracle's restatement of those lines."""
    m = curved_tet10(jf.mesh, 3, 2, 2) if et == 10 else distorted_hex8(jf.mesh, 4)
    rng = np.random.default_rng(5)
    u = (m.coords @ (0.05 * rng.standard_normal((3, 3))).T).ravel() + 1e-3 * jf.mesh.test_vector(m.n_dofs) / 1e-3 * 0.01
    v = jf.mesh.test_vector(m.n_dofs, seed=3)
    h = make(L, m, kind=L.MAT_STVK, par=LE)
    f = h.internal_force(u)
    fr = oracle.matfree(et, m.coords, m.conn, u, kind=0, par=LE, finite_strain=True)
    assert relerr(f, fr) < TOL
    h.set_linearization(u)
    for geo in (0, 1):
        h.set_option("geometric_stiffness", geo)
        rp, ci, vals, _ = oracle.assemble_csr(et, m.coords, m.conn, u=u, kind=0, par=LE, finite_strain=True, geometric=bool(geo))
        assert relerr(h.matvec(v, flags=L.TANGENT), oracle.spmv(rp, ci, vals, v)) < 1e-11
        vg, fg = h.assemble_csr(u, want_f=True)
        assert relerr(vg, vals) < 1e-11 and relerr(fg, fr) < TOL
    # with Kg the tangent is the derivative of f_int
    eps = 1e-6
    fd = (h.internal_force(u + eps * v / 1e-3) - h.internal_force(u - eps * v / 1e-3)) / (2 * eps / 1e-3)
    assert relerr(fd, h.matvec(v, flags=L.TANGENT)) < 1e-6


def test_classic_finite_strain_problem_is_stvk_and_converges(oracle, jf):
    """Problem(Elasticity) with finite_strain = true must be St. Venant-Kirchhoff (not Neo-Hookean), and run_
    must report convergence."""
    from juliafem.jl_b200 import api as A
    m = jf.mesh.tet10_kuhn(6, 2, 2, 3.0, 1.0, 1.0)
    els = [A.Element(A.Tet10, c, fields={"geometry": m.coords[c - 1].T, "youngs modulus": 1.0e7, "poissons ratio": 0.3}) for c in m.conn]
    A.update_(els, "displacement load 3", -2.0e4)
    model = A.Problem(A.Elasticity, "beam", 3)
    model.properties.finite_strain = True
    model.properties.geometric_stiffness = True
    A.add_elements_(model, els)
    left = np.nonzero(m.coords[:, 0] == 0)[0] + 1
    fixed = A.Problem(A.Dirichlet, "fixed", 3, "displacement")
    fel = [A.Element(A.Poi1, [int(n)]) for n in left]
    for c in (1, 2, 3):
        A.update_(fel, f"displacement {c}", 0.0)
    A.add_elements_(fixed, fel)
    an = A.Analysis(A.Nonlinear, model, fixed)
    A.run_(an, tol=1e-9, max_iter=20000)
    assert an.converged and an.iterations >= 2
    # the converged state satisfies the oracle's StVK equilibrium on the free dofs
    R = oracle.matfree(10, m.coords, m.conn, an.u, kind=0, par=(1.0e7, 0.3), finite_strain=True) - model._data.f_ext
    fd = model._data.fixed_dofs - 1
    R[fd] = 0
    assert np.linalg.norm(R) < 1e-7 * np.linalg.norm(model._data.f_ext)
    # ... and differs visibly from the small-strain answer (the load is large enough for the nonlinearity to matter)
    model2 = A.Problem(A.Elasticity, "beam", 3)
    A.add_elements_(model2, els)
    an2 = A.Analysis(A.Linear, model2, fixed)
    A.run_(an2, tol=1e-10)
    assert relerr(an.u, an2.u) > 1e-3
    # non-convergence is reported: one Newton step is not enough
    an3 = A.Analysis(A.Nonlinear, model, fixed)
    an3.properties.max_iterations = 1
    import juliafem.jl_b200.api as api
    with pytest.raises(api.ConvergenceError):
        api.run_(an3, tol=1e-14, max_iter=5, newton_tol=1e-14)


def test_penalty_dirichlet_on_assembled_matrix(L, oracle, jf):
    """apply_dirichlet_bc! of the CPU backend (src/element_assembly_structures.jl:237-252)."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    m = jf.mesh.tet10_kuhn(4, 2, 2, 2.0, 1.0, 1.0)
    fixed = jf.mesh.clamp_dofs(m)
    pres = np.zeros(fixed.size); pres[2::3] = 1e-3
    h = make(L, m)
    h.set_dirichlet(fixed, pres)
    vals, _ = h.assemble_csr(None)
    rp, ci = h.csr_pattern()
    K = sp.csr_matrix((vals, ci - 1, rp - 1))
    b = h.body_load((0.0, 0.0, -1.0e5))
    pen, rhs = h.csr_penalty_bc(b)
    assert abs(pen / (1e10 * np.abs(vals).max()) - 1.0) < 1e-14
    want = b.copy(); want[fixed - 1] = pen * pres
    assert np.array_equal(rhs, want)
    # the device matrix now carries the penalty on the fixed diagonal entries: K_pen e_i = K e_i + pen e_i
    i = int(fixed[5] - 1)
    e = np.zeros(m.n_dofs); e[i] = 1.0
    y = h.spmv(e)
    assert abs(y[i] - (K[i, i] + pen)) <= 1e-15 * pen
    # solving the penalised system reproduces the eliminated solution to ~1/penalty
    Kp = K + sp.csr_matrix((np.full(fixed.size, pen), (fixed - 1, fixed - 1)), shape=K.shape)
    up = spla.spsolve(Kp.tocsc(), rhs)
    free = np.setdiff1d(np.arange(m.n_dofs), fixed - 1)
    ub = np.zeros(m.n_dofs); ub[fixed - 1] = pres
    ue = ub.copy()
    ue[free] = spla.spsolve(K[free][:, free].tocsc(), (b - K @ ub)[free])
    assert relerr(up, ue) < 1e-6
