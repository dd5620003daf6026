"""GPU: the CUDA path (through the C ABI of include/jfem_b200.h) against the CPU oracle on the same inputs.

Tolerances (BASELINE.json north_star): CSR pattern / DOF numbering bit-exact; stiffness entries and K.u within
1e-12 relative (max-norm relative to the largest entry); CG solutions within 1e-8 relative residual."""
import os

import numpy as np
import pytest

from conftest import curved_tet10, distorted_hex8, relerr

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
TOL = 1e-12
LE = (210e9, 0.3)          # demos/cantilever_physics_gpu.jl:58-59
NH = (3e6, 0.45)           # src/materials/neo_hookean.jl:60
PP = (200e9, 0.3, 250e6, 1e9)   # docs/book/perfect_plasticity_implementation.md:277-282


@pytest.fixture(scope="module")
def L(jf):
    from juliafem.jl_b200 import _lib
    assert _lib.device_count() > 0, "no GPU visible"
    return _lib


def make(L, m, kind=0, par=LE, **opts):
    h = L.Handle(m.elem_type, m.coords, m.conn)
    for k, v in opts.items():
        h.set_option(k, v)
    h.set_material(kind, par)
    return h


def fixture_mesh(jf):
    d = np.load(os.path.join(HERE, "golden", "tet10_fixture.npz"))
    return jf.mesh.Mesh(10, d["coords"], d["conn"])


# ------------------------------------------------------------------------------------------------ K.u

@pytest.mark.parametrize("patch", [128, 256, 512])
@pytest.mark.parametrize("affine", [0, 1])
def test_matvec_tet10_structured(L, oracle, jf, patch, affine):
    m = jf.mesh.tet10_kuhn(8, 4, 4, 4.0, 1.0, 1.0)
    u = jf.mesh.test_vector(m.n_dofs)
    ref = oracle.matfree(10, m.coords, m.conn, u, par=LE)
    h = make(L, m, patch_elems=patch, affine_fast_path=affine)
    y = h.matvec(u)
    info = h.info()
    assert info.n_affine_elems == (m.n_elems if affine else 0)
    assert relerr(y, ref) < TOL
    # deterministic: bitwise identical on repeat
    assert np.array_equal(h.matvec(u), y)


def test_matvec_tet10_curved(L, oracle, jf):
    m = curved_tet10(jf.mesh)
    u = jf.mesh.test_vector(m.n_dofs)
    h = make(L, m)
    assert relerr(h.matvec(u), oracle.matfree(10, m.coords, m.conn, u, par=LE)) < TOL
    assert h.info().n_affine_elems == 0


def test_matvec_tet10_mixed_affine_and_curved(L, oracle, jf):
    m = jf.mesh.tet10_kuhn(6, 3, 3)
    rng = np.random.default_rng(5)
    sel = m.coords[:, 0] > 0.5
    m.coords[sel] += 0.01 * rng.standard_normal((int(sel.sum()), 3)) / 12
    u = jf.mesh.test_vector(m.n_dofs)
    h = make(L, m)
    assert relerr(h.matvec(u), oracle.matfree(10, m.coords, m.conn, u, par=LE)) < TOL
    assert 0 < h.info().n_affine_elems < m.n_elems          # two launches: affine closed form + isoparametric


def test_matvec_reference_fixture_mesh(L, oracle, jf):
    m = fixture_mesh(jf)
    u = jf.mesh.test_vector(m.n_dofs)
    h = make(L, m)
    assert relerr(h.matvec(u), oracle.matfree(10, m.coords, m.conn, u, par=LE)) < TOL


def test_matvec_hex8(L, oracle, jf):
    for m in (jf.mesh.hex8_lattice(9, 7, 6, 0.25), distorted_hex8(jf.mesh)):
        u = jf.mesh.test_vector(m.n_dofs)
        h = make(L, m)
        assert relerr(h.matvec(u), oracle.matfree(8, m.coords, m.conn, u, par=LE)) < TOL


def test_matvec_tet4(L, oracle, jf):
    m = jf.mesh.tet4_kuhn(5, 4, 3)
    u = jf.mesh.test_vector(m.n_dofs)
    h = make(L, m)
    assert relerr(h.matvec(u), oracle.matfree(4, m.coords, m.conn, u, par=LE)) < TOL


def test_matvec_projection_rigid_modes_linearity(L, oracle, jf):
    m = jf.mesh.tet10_kuhn(6, 2, 2, 3.0, 1.0, 1.0)
    fixed = jf.mesh.clamp_dofs(m)
    u = jf.mesh.test_vector(m.n_dofs, fixed)
    h = make(L, m)
    h.set_dirichlet(fixed)
    assert h.info().n_fixed == fixed.size
    ref = oracle.matfree(10, m.coords, m.conn, u, par=LE, fixed_dofs=fixed)
    y = h.matvec(u, flags=L.PROJECT)
    assert relerr(y, ref) < TOL and np.all(y[fixed - 1] == 0)
    # pure K.v keeps the reaction rows
    assert np.abs(h.matvec(u)[fixed - 1]).max() > 0
    # rigid translation + infinitesimal rotation are in the kernel of K
    W = np.array([[0, -1, 2], [1, 0, -3], [-2, 3, 0]], float) * 1e-3
    rb = (m.coords @ W.T + np.array([1e-3, -2e-3, 5e-4])).ravel()
    scale = np.abs(h.matvec(u)).max() / np.abs(u).max()
    assert np.abs(h.matvec(rb)).max() < 1e-11 * scale * np.abs(rb).max()
    # linearity
    v = jf.mesh.test_vector(m.n_dofs, seed=99)
    lhs = h.matvec(2.5 * u - 0.5 * v)
    rhs = 2.5 * h.matvec(u) - 0.5 * h.matvec(v)
    assert relerr(lhs, rhs) < 1e-13
    # symmetry: v.Ku == u.Kv
    a, b = v @ h.matvec(u), u @ h.matvec(v)
    assert abs(a - b) < 1e-12 * abs(a)


def test_matvec_atomic_mode_and_orphan_nodes(L, oracle, jf):
    m = jf.mesh.tet10_kuhn(5, 3, 3)
    # add two nodes that no element references: y must be 0 there
    m.coords = np.vstack([m.coords, [[9, 9, 9], [8, 8, 8]]])
    u = jf.mesh.test_vector(m.n_dofs)
    ref = oracle.matfree(10, m.coords, m.conn, u, par=LE)
    h = make(L, m)
    y = h.matvec(u)
    assert relerr(y, ref) < TOL and np.all(y[-6:] == 0)
    h2 = make(L, m, deterministic=0)
    assert relerr(h2.matvec(u), ref) < TOL


@pytest.mark.parametrize("opts", [dict(warp_specialised=0), dict(lane_window=0), dict(fused_interface=1), dict(lane_window=256, patch_elems=256)])
def test_matvec_kernel_variants_agree(L, oracle, jf, opts):
    """The generic patch kernel, the unoptimised lane order and the cooperative-launch interface reduction are the same
    operator as the default warp-specialised kernel: each within 1e-12 of the oracle and deterministic."""
    m = jf.mesh.tet10_kuhn(12, 6, 5, 3.0, 1.0, 1.0)          # 2160 elements: 9 patches, one partial
    fixed = jf.mesh.clamp_dofs(m)
    u = jf.mesh.test_vector(m.n_dofs, fixed)
    ref = oracle.matfree(10, m.coords, m.conn, u, par=LE, fixed_dofs=fixed)
    h = make(L, m, **opts)
    h.set_dirichlet(fixed)
    y = h.matvec(u, flags=L.PROJECT)
    assert relerr(y, ref) < TOL
    for _ in range(3):
        assert np.array_equal(h.matvec(u, flags=L.PROJECT), y)
    h0 = make(L, m)
    h0.set_dirichlet(fixed)
    assert relerr(h0.matvec(u, flags=L.PROJECT), y) < 1e-13


def test_matvec_many_patches_per_sm_pipeline(L, oracle, jf):
    """More patches than SMs x 2: every block runs the steady-state pipeline (blob rings wrap, both staging tiles reused)."""
    m = jf.mesh.tet10_kuhn(40, 12, 12, 4.0, 1.0, 1.0)         # 34 560 elements = 135 patches on 148 SMs ... use EP=128 for 270
    u = jf.mesh.test_vector(m.n_dofs)
    ref = oracle.matfree(10, m.coords, m.conn, u, par=LE)
    for ep in (256, 128):
        h = make(L, m, patch_elems=ep)
        assert relerr(h.matvec(u), ref) < TOL
    m = jf.mesh.tet10_kuhn(96, 24, 24, 4.0, 1.0, 1.0)         # 331 776 elements = 1296 patches: ~9 per SM
    u = jf.mesh.test_vector(m.n_dofs)
    h = make(L, m)
    y = h.matvec(u)
    ref = oracle.matfree(10, m.coords, m.conn, u, par=LE)
    assert relerr(y, ref) < TOL


def test_matvec_device_pointers(L, oracle, jf):
    import torch
    m = jf.mesh.tet10_kuhn(8, 4, 4)
    u = jf.mesh.test_vector(m.n_dofs)
    h = make(L, m)
    xd = torch.from_numpy(u).cuda()
    yd = torch.full_like(xd, float("nan"))
    h.set_stream(torch.cuda.current_stream().cuda_stream)
    h.matvec(xd, yd)
    torch.cuda.synchronize()
    assert relerr(yd.cpu().numpy(), oracle.matfree(10, m.coords, m.conn, u, par=LE)) < TOL
    assert h.info().matvec_launches == 2


def test_matvec_large_size_properties(L, jf):
    """BASELINE config[1] size (T1: 88x22x22 cells, 1 075 275 DOF): size-independent properties only."""
    m = jf.mesh.tet10_kuhn(88, 22, 22, 4.0, 1.0, 1.0)
    assert m.n_dofs == 1075275 and m.n_elems == 255552
    h = make(L, m)
    u = jf.mesh.test_vector(m.n_dofs)
    v = jf.mesh.test_vector(m.n_dofs, seed=4)
    Ku, Kv = h.matvec(u), h.matvec(v)
    assert abs(v @ Ku - u @ Kv) < 1e-11 * abs(v @ Ku)                      # symmetry
    assert u @ Ku > 0                                                       # positive semi-definite
    t = np.tile([1e-3, -2e-3, 5e-4], m.n_nodes)
    assert np.abs(h.matvec(t)).max() < 1e-10 * np.abs(Ku).max()            # translations in the kernel
    assert np.abs(Ku.reshape(-1, 3).sum(axis=0)).max() < 1e-9 * np.abs(Ku).max()   # sum of internal forces = 0
    h0 = make(L, m, affine_fast_path=0, patch_elems=128)
    assert relerr(h0.matvec(u), Ku) < TOL                                   # general path == affine closed form


# ------------------------------------------------------------------------------------------------ element matrices / CSR

@pytest.mark.parametrize("et", [4, 8, 10])
def test_element_matrices_linear(L, oracle, jf, et):
    m = {4: lambda: jf.mesh.tet4_kuhn(2, 2, 2), 8: lambda: distorted_hex8(jf.mesh, 4), 10: lambda: curved_tet10(jf.mesh, 2, 2, 2)}[et]()
    h = make(L, m)
    K, f = h.element_matrices()
    for e in range(0, m.n_elems, max(1, m.n_elems // 25)):
        Ke, _, _, _ = oracle.element(et, m.coords[m.conn[e] - 1], par=LE)
        assert relerr(K[e], Ke) < TOL
    assert np.all(f == 0)


def test_csr_pattern_bit_exact_and_values(L, oracle, jf):
    for m in (jf.mesh.tet10_kuhn(5, 3, 2), fixture_mesh(jf), distorted_hex8(jf.mesh, 5)):
        h = L.Handle(m.elem_type, m.coords, m.conn, index_base=1)
        h.set_material(0, LE)
        rowptr, colind = h.csr_pattern()
        rp, ci, vals_ref, _ = oracle.assemble_csr(m.elem_type, m.coords, m.conn, par=LE)
        assert rowptr.dtype == np.int64 and colind.dtype == np.int32
        assert np.array_equal(rowptr, rp + 1) and np.array_equal(colind, ci + 1)    # 1-based like colptr/rowval
        vals, _ = h.assemble_csr()
        assert relerr(vals, vals_ref) < TOL
        assert np.array_equal(vals != 0, vals != 0)
        u = jf.mesh.test_vector(m.n_dofs)
        assert relerr(h.spmv(u), oracle.spmv(rp, ci, vals_ref, u)) < TOL
        assert relerr(h.spmv(u), h.matvec(u)) < TOL
        # symmetrised variant (src/solvers.jl:289-292)
        vs, _ = h.assemble_csr(symmetrise=True)
        _, _, vsr, _ = oracle.assemble_csr(m.elem_type, m.coords, m.conn, par=LE, symmetrise=True)
        assert relerr(vs, vsr) < TOL


@pytest.mark.parametrize("kernel", [0, 1])
def test_assembly_kernels_against_oracle(L, oracle, jf, kernel):
    """Both assembly kernels (option assembly_kernel: 1 = one warp per element with shared Gauss-point geometry [default],
    0 = one thread per (element, column)) against the oracle: dense Ke of every element type, CSR values for the
    linear-elastic, Neo-Hookean (finite strain + geometric stiffness) and plastic tangents, per-element parameters."""
    for et, m in ((4, jf.mesh.tet4_kuhn(3, 2, 2)), (8, distorted_hex8(jf.mesh, 4)), (10, curved_tet10(jf.mesh, 3, 2, 2))):
        h = make(L, m, assembly_kernel=kernel)
        K, _ = h.element_matrices()
        for e in range(0, m.n_elems, max(1, m.n_elems // 40)):
            Ke, _, _, _ = oracle.element(et, m.coords[m.conn[e] - 1], par=LE)
            assert relerr(K[e], Ke) < TOL
        vals, _ = h.assemble_csr()
        _, _, vr, _ = oracle.assemble_csr(et, m.coords, m.conn, par=LE)
        assert relerr(vals, vr) < TOL
        h.close()
    m = curved_tet10(jf.mesh, 4, 3, 3)
    rng = np.random.default_rng(11)
    u = 0.01 * rng.standard_normal(m.n_dofs) * np.abs(m.coords).max()
    h = make(L, m, 1, NH, assembly_kernel=kernel)
    vals, _ = h.assemble_csr(u)
    _, _, vr, _ = oracle.assemble_csr(10, m.coords, m.conn, u=u, kind=1, par=NH, finite_strain=True, geometric=True)
    assert relerr(vals, vr) < 1e-11
    h.close()
    # plastic tangent at a yielded trial state: first step commits a state, second assembles around it
    pp = (200e9, 0.3, 100e6, 10e9)
    u1 = 2e-3 * rng.standard_normal(m.n_dofs) * 0.1
    h = make(L, m, 2, pp, assembly_kernel=kernel)
    h.internal_force(u1)
    h.commit_state()
    _, _, _, _, s1 = oracle.assemble_csr(10, m.coords, m.conn, u=u1, kind=2, par=pp, want_state=True)
    assert (s1[:, :, 12] > 0).mean() > 0.2
    u2 = u1 + 5e-5 * rng.standard_normal(m.n_dofs)
    vals, _ = h.assemble_csr(u2)
    _, _, vr, _ = oracle.assemble_csr(10, m.coords, m.conn, u=u2, kind=2, par=pp, state_old=s1)
    assert relerr(vals, vr) < 1e-11
    h.close()
    # per-element Young's modulus / Poisson's ratio (E_vec / nu_vec of ext:135-141)
    m = jf.mesh.tet10_kuhn(3, 2, 2)
    par = np.stack([210e9 * (1.0 + 0.5 * rng.random(m.n_elems)), 0.2 + 0.2 * rng.random(m.n_elems)], axis=1)
    h = L.Handle(10, m.coords, m.conn)
    h.set_option("assembly_kernel", kernel)
    h.set_material(0, par)
    K, _ = h.element_matrices()
    for e in range(0, m.n_elems, 7):
        Ke, _, _, _ = oracle.element(10, m.coords[m.conn[e] - 1], par=tuple(par[e]))
        assert relerr(K[e], Ke) < TOL
    h.close()


# ------------------------------------------------------------------------------------------------ CG

def test_cg_matches_reference_iteration(L, oracle, jf):
    m = jf.mesh.tet10_kuhn(8, 2, 2, 4.0, 1.0, 1.0)
    fixed = jf.mesh.clamp_dofs(m)
    b = np.zeros(m.n_dofs); b[2::3] = -1e3
    rp, ci, vals, _ = oracle.assemble_csr(10, m.coords, m.conn, par=LE)
    h = make(L, m)
    h.set_dirichlet(fixed)
    # few iterations, absolute tolerance (reference semantics, src/backend/cpu.jl:244): iterates must agree closely
    xr, itr, rr = oracle.cg_csr(rp, ci, vals, b, fixed_dofs=fixed, tol=1e-30, max_iter=25)
    x, it, res = h.cg(b, tol=1e-30, max_iter=25)
    assert it == itr == 25
    assert relerr(x, xr) < 1e-9 and abs(res - rr) < 1e-8 * rr
    # converge: relative 1e-8 (north_star) -> true residual check
    x, it, res = h.cg(b, tol=1e-8, relative=True, max_iter=20000)
    xr, itr, _ = oracle.cg_csr(rp, ci, vals, b, fixed_dofs=fixed, tol=1e-8, relative=True, max_iter=20000)
    assert it < 20000 and abs(it - itr) <= max(3, itr // 50)
    r = b - oracle.spmv(rp, ci, vals, x)
    r[fixed - 1] = 0
    bn = b.copy(); bn[fixed - 1] = 0
    assert np.linalg.norm(r) <= 2e-8 * np.linalg.norm(bn)
    assert np.all(x[fixed - 1] == 0)
    assert relerr(x, xr) < 1e-6
    # assembled operator route gives the same answer
    h.assemble_csr()
    x2, it2, _ = h.cg(b, tol=1e-8, relative=True, max_iter=20000, flags=L.USE_CSR)
    assert relerr(x2, x) < 1e-6 and abs(it2 - it) <= max(3, it // 50)
    # max_iter reached is not an error (ext/JuliaFEMCUDAExt.jl:630-631)
    _, it3, _ = h.cg(b, tol=1e-30, max_iter=3)
    assert it3 == 3
    # zero right-hand side: early exit with 0 iterations (src/backend/cpu.jl:230-233)
    _, it4, r4 = h.cg(np.zeros(m.n_dofs), tol=1e-6)
    assert it4 == 0 and r4 == 0.0


# ------------------------------------------------------------------------------------------------ nonlinear materials

def _nh_u(jf, m, amp):
    rng = np.random.default_rng(11)
    G = amp * rng.standard_normal((3, 3))
    return (m.coords @ G.T).ravel() + 0.1 * amp * jf.mesh.test_vector(m.n_dofs) / 1e-3 * (1.0 / 12)


@pytest.mark.parametrize("et", [10, 8, 4])
def test_neo_hookean_internal_force_and_tangent(L, oracle, jf, et):
    m = {10: lambda: curved_tet10(jf.mesh, 3, 2, 2), 8: lambda: distorted_hex8(jf.mesh, 4), 4: lambda: jf.mesh.tet4_kuhn(3, 2, 2)}[et]()
    u = _nh_u(jf, m, 0.1)
    h = make(L, m, kind=1, par=NH)
    f = h.internal_force(u)
    fr = oracle.matfree(et, m.coords, m.conn, u, kind=1, par=NH, finite_strain=True)
    assert relerr(f, fr) < TOL
    # tangent (Km + Kg) action vs assembled oracle tangent
    rp, ci, vals, _ = oracle.assemble_csr(et, m.coords, m.conn, u=u, kind=1, par=NH, finite_strain=True, geometric=True)
    v = jf.mesh.test_vector(m.n_dofs, seed=3)
    h.set_linearization(u)
    assert relerr(h.matvec(v, flags=L.TANGENT), oracle.spmv(rp, ci, vals, v)) < 1e-11
    # element matrices and assembled CSR
    K, fe = h.element_matrices(u)
    for e in range(0, m.n_elems, max(1, m.n_elems // 10)):
        Km, Kg, fo, _ = oracle.element(et, m.coords[m.conn[e] - 1], u.reshape(-1, 3)[m.conn[e] - 1], kind=1, par=NH, finite_strain=True, geometric=True)
        assert relerr(K[e], Km + Kg) < 1e-11 and relerr(fe[e], fo) < TOL
    vg, fg = h.assemble_csr(u, want_f=True)
    assert relerr(vg, vals) < 1e-11 and relerr(fg, fr) < TOL
    # tangent is the derivative of the internal force
    eps = 1e-6
    fd = (h.internal_force(u + eps * v / 1e-3) - h.internal_force(u - eps * v / 1e-3)) / (2 * eps / 1e-3)
    assert relerr(fd, h.matvec(v, flags=L.TANGENT)) < 1e-6


def test_neo_hookean_invalid_deformation_raises(L, jf):
    m = jf.mesh.tet10_kuhn(2, 2, 2)
    h = make(L, m, kind=1, par=NH)
    u = (m.coords @ (-2.0 * np.eye(3)).T).ravel()     # F = -I  -> det C > 0 but collapse through zero: use F with det<=0
    u = (m.coords @ np.diag([-1.0, 0.0, 0.0]).T).ravel()   # F = diag(0,1,1) -> det C = 0
    with pytest.raises(L.DomainError):
        h.internal_force(u)


def test_perfect_plasticity_state_update(L, oracle, jf):
    m = curved_tet10(jf.mesh, 3, 2, 2, amp=0.01)
    G = np.array([[3e-3, 1e-3, 0], [0, -1e-3, 5e-4], [2e-4, 0, -8e-4]])
    u = (m.coords @ G.T).ravel() + 0.3 * jf.mesh.test_vector(m.n_dofs)
    h = make(L, m, kind=2, par=PP)
    f = h.internal_force(u)
    rp, ci, vals, fr, sn = oracle.assemble_csr(10, m.coords, m.conn, u=u, kind=2, par=PP, want_state=True)
    assert relerr(f, fr) < TOL
    trial = h.get_state(committed=False)
    assert np.all(h.get_state(committed=True) == 0)            # not stored until commit (abstract_material.jl:203-207)
    assert relerr(trial, sn) < 1e-11
    frac = np.mean(sn[:, :, 12] > 0)
    assert 0.05 < frac <= 1.0
    # tangent from the committed (old) state at u
    v = jf.mesh.test_vector(m.n_dofs, seed=8)
    h.set_linearization(u)
    assert relerr(h.matvec(v, flags=L.TANGENT), oracle.spmv(rp, ci, vals, v)) < 1e-11
    vg, _ = h.assemble_csr(u)
    assert relerr(vg, vals) < 1e-11
    # commit, then a second step from the new state
    h.commit_state()
    assert relerr(h.get_state(committed=True), sn) < 1e-11
    u2 = 1.5 * u
    f2 = h.internal_force(u2)
    _, _, _, fr2, sn2 = oracle.assemble_csr(10, m.coords, m.conn, u=u2, kind=2, par=PP, state_old=sn, want_state=True)
    assert relerr(f2, fr2) < TOL and relerr(h.get_state(committed=False), sn2) < 1e-11
    # set_state round trip
    h.set_state(sn2)
    assert np.array_equal(h.get_state(), sn2)


def test_newton_krylov_neo_hookean(L, oracle, jf):
    m = jf.mesh.tet10_kuhn(6, 2, 2, 3.0, 1.0, 1.0)
    fixed = jf.mesh.clamp_dofs(m)
    h = make(L, m, kind=1, par=NH)
    h.set_dirichlet(fixed)
    top = np.nonzero(np.abs(m.coords[:, 2] - 1.0) < 1e-12)[0]
    fext = np.zeros(m.n_dofs)
    fext[3 * top + 2] = -6e3 / top.size          # ~0.1 tip deflection on a 3 x 1 x 1 beam with E = 3e6: clearly nonlinear
    # the reference's forcing term eta = min(forcing_max, ||R||^0.5) is norm-unit dependent; with forces of O(1e3) its
    # default forcing_max = 0.9 would only ask for a 10 % reduction per linear solve, so tighten it for this test
    u, nit, cgit, res, hist = h.newton_krylov(fext, newton_tol=1e-6, max_newton=30, max_cg_per_newton=4000, forcing_max=1e-3)
    assert res < 1e-6 and 1 < nit <= 30 and len(hist) == nit and cgit == sum(c for c, _, _ in hist)
    assert all(abs(eta - min(1e-3, rn ** 0.5)) < 1e-12 for _, rn, eta in hist)     # ext/JuliaFEMCUDAExt.jl:819
    R = fext - oracle.matfree(10, m.coords, m.conn, u, kind=1, par=NH, finite_strain=True)
    R[fixed - 1] = 0
    assert np.linalg.norm(R) < 1e-5
    assert np.abs(u).max() > 1e-3 and np.all(u[fixed - 1] == 0)


def test_block_jacobi_pcg(L, oracle, jf):
    """Opt-in 3x3 block-Jacobi PCG: same solution as plain CG, diagonal blocks equal the assembled ones."""
    m = curved_tet10(jf.mesh, 8, 2, 2, amp=0.01)
    fixed = jf.mesh.clamp_dofs(m, tol=0.02)
    b = np.zeros(m.n_dofs); b[2::3] = -1e3
    h = make(L, m)
    h.set_dirichlet(fixed)
    x0, it0, _ = h.cg(b, tol=1e-9, relative=True, max_iter=50000)
    x1, it1, r1 = h.cg(b, tol=1e-9, relative=True, max_iter=50000, flags=L.JACOBI)
    assert it1 < 50000 and it1 <= it0
    assert relerr(x1, x0) < 1e-6 and np.all(x1[fixed - 1] == 0)
    rp, ci, vals, _ = oracle.assemble_csr(10, m.coords, m.conn, par=LE)
    r = b - oracle.spmv(rp, ci, vals, x1); r[fixed - 1] = 0
    bn = b.copy(); bn[fixed - 1] = 0
    assert np.linalg.norm(r) <= 5e-9 * np.linalg.norm(bn)


def test_per_element_material(L, oracle, jf):
    """Per-element E, nu (E_vec / nu_vec of ext/JuliaFEMCUDAExt.jl:135-141): affine kernel, curved kernel, assembled path."""
    for m in (jf.mesh.tet10_kuhn(4, 2, 2), curved_tet10(jf.mesh, 3, 2, 2)):
        rng = np.random.default_rng(3)
        par = np.stack([210e9 * (0.5 + rng.random(m.n_elems)), 0.2 + 0.2 * rng.random(m.n_elems)], axis=1)
        u = jf.mesh.test_vector(m.n_dofs)
        ref = np.zeros(m.n_dofs)
        for e in range(m.n_elems):
            Ke, _, _, _ = oracle.element(10, m.coords[m.conn[e] - 1], par=tuple(par[e]))
            g = (3 * (m.conn[e][:, None] - 1) + np.arange(3)[None, :]).ravel()
            ref[g] += Ke @ u[g]
        h = L.Handle(10, m.coords, m.conn)
        h.set_material(0, par)
        assert relerr(h.matvec(u), ref) < TOL
        K, _ = h.element_matrices()
        Ke, _, _, _ = oracle.element(10, m.coords[m.conn[5] - 1], par=tuple(par[5]))
        assert relerr(K[5], Ke) < TOL
        h.assemble_csr()
        assert relerr(h.spmv(u), ref) < TOL
        # switching back to a homogeneous material works
        h.set_material(0, LE)
        assert relerr(h.matvec(u), oracle.matfree(10, m.coords, m.conn, u, par=LE)) < TOL
    with pytest.raises(L.JfemError):
        h.set_material(0, np.array([[1.0, 0.3]] * (m.n_elems - 1) + [[-1.0, 0.3]]))
