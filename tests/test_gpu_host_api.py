"""GPU: the reference-facing host interface (Physics/solve_ and Problem/Analysis/run_) against the oracle."""
import numpy as np
import pytest

from conftest import relerr

pytestmark = pytest.mark.gpu


def test_library_is_native_and_loaded(jf):
    from juliafem.jl_b200 import _lib
    assert _lib.lib().jfem_abi_version() == 2
    assert "libjfem_b200.so" in open("/proc/self/maps").read()


def _cantilever(jf, A, topo="Tet10"):
    m = jf.mesh.tet10_kuhn(6, 2, 2, 3.0, 1.0, 1.0)
    # arbitrary (non-dense) node ids, as the reference allows (ext/JuliaFEMCUDAExt.jl:100-108)
    ids = 7 * np.arange(1, m.n_nodes + 1) + 3
    els = [A.Element(A.Tet10, ids[c - 1], fields={"geometry": m.coords[c - 1].T, "youngs_modulus": 210e9, "poissons_ratio": 0.3})
           for c in m.conn]
    return m, ids, els


def _direct_reference(oracle, m, fixed, b, prescribed=None):
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    rp, ci, vals, _ = oracle.assemble_csr(10, m.coords, m.conn, par=(210e9, 0.3))
    K = sp.csr_matrix((vals, ci, rp))
    u = np.zeros(m.n_dofs) if prescribed is None else prescribed.copy()
    free = np.setdiff1d(np.arange(m.n_dofs), fixed - 1)
    rhs = b - K @ u
    u[free] += spla.spsolve(K[free][:, free].tocsc(), rhs[free])     # src/solvers.jl:205-210
    return u


def test_physics_solve_gpu_backend(oracle, jf):
    from juliafem.jl_b200 import api as A
    m, ids, els = _cantilever(jf, A)
    physics = A.Physics(A.Elasticity, "cantilever beam", 3)
    A.add_elements_(physics, els)
    left = np.nonzero(m.coords[:, 0] == 0)[0]
    for n in left:
        A.add_dirichlet_(physics, [ids[n]], [1, 2, 3], 0.0)
    # pressure -1e6 on the top face through lumped Tri3 facets (ext/JuliaFEMCUDAExt.jl:368-416)
    top = np.nonzero(np.abs(m.coords[:, 2] - 1.0) < 1e-12)[0]
    from scipy.spatial import Delaunay
    tri = Delaunay(m.coords[top][:, :2]).simplices
    area = 0.0
    for t in tri:
        nodes = top[t]
        A.add_neumann_(physics, A.Element(A.Tri3, ids[nodes], fields={"geometry": m.coords[nodes].T}), (0.0, 0.0, -1e6))
    res = A.solve_(physics, backend=A.GPU(), tol=1e-3, max_iter=20000)
    assert isinstance(res, A.ElasticitySolution) and res.newton_iterations == 1 and res.cg_iterations < 20000
    assert res.residual < 1e-3 and len(res.history) == 1
    # reference answer: same lumped load, elimination + direct solve on the oracle matrix
    b = np.zeros(m.n_dofs)
    for t in tri:
        X = m.coords[top[t]]
        a = 0.5 * np.linalg.norm(np.cross(X[1] - X[0], X[2] - X[0]))
        for n in top[t]:
            b[3 * n + 2] += -1e6 * a / 3
    assert abs(b.sum() + 3.0e6) < 1e-3        # total force = pressure * area
    fixed = (3 * left[:, None] + np.arange(1, 4)).ravel()
    uref = _direct_reference(oracle, m, fixed, b)
    assert relerr(res.u, uref) < 1e-6
    assert np.all(res.u[fixed - 1] == 0)


def test_classic_problem_analysis_run(oracle, jf):
    from juliafem.jl_b200 import api as A
    m, ids, els = _cantilever(jf, A)
    model = A.Problem(A.Elasticity, "body", 3)
    for el in els:
        el.fields = {"geometry": el.fields["geometry"]}
    A.update_(els, "youngs modulus", 210e9)
    A.update_(els, "poissons ratio", 0.3)
    A.update_(els, "displacement load 3", -7.8e4)          # body load, examples/linear_static.jl:85
    A.add_elements_(model, els)
    fixed = A.Problem(A.Dirichlet, "fixed", 3, "displacement")
    left = np.nonzero(m.coords[:, 0] == 0)[0]
    fel = [A.Element(A.Poi1, [ids[n]]) for n in left]      # create_nodal_elements (examples/linear_static.jl:76)
    A.update_(fel, "displacement 1", 0.0)
    A.update_(fel, "displacement 2", 0.0)
    A.update_(fel, "displacement 3", 1e-3)                  # non-homogeneous: lifted into the right-hand side
    A.add_elements_(fixed, fel)
    # assemble!: K and f on the reference's pattern
    A.assemble_(model, 0.0)
    rp, ci, vals, _ = oracle.assemble_csr(10, m.coords, m.conn, par=(210e9, 0.3))
    K = model.assembly.K
    assert np.array_equal(K.rowptr, rp + 1) and np.array_equal(K.colind, ci + 1) and relerr(K.vals, vals) < 1e-12
    I, J, V = K.to_coo()
    assert I.size == vals.size and I.min() == 1
    # consistent body load: sum = b * volume
    assert abs(model.assembly.f[2::3].sum() + 7.8e4 * 3.0) < 1e-6
    analysis = A.Analysis(A.Linear, model, fixed)
    A.run_(analysis)
    fd = (3 * left[:, None] + np.arange(1, 4)).ravel()
    pres = np.zeros(m.n_dofs); pres[3 * left + 2] = 1e-3
    uref = _direct_reference(oracle, m, fd, model.assembly.f, pres)
    assert relerr(analysis.u, uref) < 1e-6 and np.allclose(analysis.u[3 * left + 2], 1e-3)
    d = A.nodal_displacements(model, analysis.u)
    assert set(d) == set(int(i) for i in ids)


def test_unsupported_element_refused_like_reference(jf):
    """test/test_problems_elasticity_assemble_3d_seg3.jl:7-12: assemble! throws for a Seg3 in a 3D problem."""
    from juliafem.jl_b200 import api as A
    p = A.Problem(A.Elasticity, "body", 3)
    with pytest.raises(ValueError):
        p.elements.append(A.Element("Seg3", [1, 2, 3]))
    el = A.Element(A.Tri3, [1, 2, 3], fields={"geometry": np.eye(3), "youngs modulus": 1.0, "poissons ratio": 0.3})
    p.elements.append(el)
    with pytest.raises(ValueError):
        A.assemble_(p, 0.0)


def test_linear_static_example_end_to_end(oracle, jf):
    """X1 = BASELINE.json configs[0]: examples/linear_static.jl through the reference-facing interface on the GPU.
    Mesh = the committed copy of JuliaFEMSMP18.med (tests/golden/make_fixtures.py); expected max |u| =
    2.4052929896922337 (examples/linear_static.jl:133; the reference tests it with isapprox, rtol 1.5e-8)."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    from juliafem.jl_b200 import api as A
    from test_oracle_pins import _linear_static_case
    m, fixed_dofs, pin = _linear_static_case()
    els = [A.Element(A.Tet10, c, fields={"geometry": m.coords[c - 1].T}) for c in m.conn]
    model = A.Problem(A.Elasticity, "OTHER", 3)
    A.update_(els, "youngs modulus", 208.0e3)                  # examples/linear_static.jl:28-30
    A.update_(els, "poissons ratio", 0.30)
    A.update_(els, "density", 7.80e-9)
    A.add_elements_(model, els)
    A.update_(els, "displacement load 1", 1.0)                 # :85
    nodes = np.union1d(jf.mesh.nodes_at_plane(m, 1, 50.0), jf.mesh.find_nearest_nodes(m, [165.0, 88.0, 10], 3))   # :46-72
    fixed = A.Problem(A.Dirichlet, "fixed", 3, "displacement")
    fel = [A.Element(A.Poi1, [int(n)]) for n in nodes]
    for c in (1, 2, 3):
        A.update_(fel, f"displacement {c}", 0.0)
    A.add_elements_(fixed, fel)
    analysis = A.Analysis(A.Linear, model, fixed)
    A.run_(analysis, tol=1e-12, relative=True, max_iter=200000)
    assert analysis.converged and analysis.residual <= 1e-12 * np.linalg.norm(model._data.f_ext[np.setdiff1d(np.arange(m.n_dofs), fixed_dofs - 1)]) * 1.01
    u = analysis("displacement", 0.0)                          # :131: Dict node id -> displacement vector
    umax = max(np.linalg.norm(v) for v in u.values())
    assert abs(umax / pin["max_u_norm"] - 1.0) < 1e-8, umax
    # and the whole field against the oracle's eliminated direct solve
    rp, ci, vals, _ = oracle.assemble_csr(10, m.coords, m.conn, par=(pin["E"], pin["nu"]), symmetrise=True)
    K = sp.csr_matrix((vals, ci, rp))
    f = oracle.body_load(10, m.coords, m.conn, (1.0, 0.0, 0.0))
    free = np.setdiff1d(np.arange(m.n_dofs), fixed_dofs - 1)
    uref = np.zeros(m.n_dofs)
    uref[free] = spla.splu(K[free][:, free].tocsc()).solve(f[free])
    assert relerr(analysis.u, uref) < 1e-8
