"""CPU-side validation of the patch tables (patches.cpp) and of the element code shared with the kernels.

tests/hostcheck/hostcheck.cu replays the kernel's data flow on the host (gather in p order -> element threads through
the element table -> jagged staging tile -> per-node ordered reduction -> interface slots summed by the last arriver)
and must reproduce the oracle's K.u.  No GPU is involved; the CUDA kernels themselves are covered by the -m gpu tests.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import curved_tet10, distorted_hex8, relerr

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def hostcheck():
    subprocess.check_call(["make", "-C", os.path.join(HERE, "hostcheck"), "-s"], stderr=subprocess.DEVNULL)
    L = C.CDLL(os.path.join(HERE, "hostcheck", "libhostcheck.so"))
    L.hostcheck_error.restype = C.c_char_p
    L.hostcheck_matvec.restype = C.c_int
    L.hostcheck_matvec.argtypes = [C.c_int, C.c_longlong, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                   C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_longlong]
    return L


def run(L, m, x, fixed=None, EP=256, window=48, affine=1, project=0, E=210e9, nu=0.3, n_owned=-1):
    coords = np.ascontiguousarray(m.coords, dtype=np.float64)
    conn = np.ascontiguousarray(m.conn - 1, dtype=np.int32)
    fx = np.zeros(m.n_dofs, dtype=np.uint8)
    if fixed is not None:
        fx[np.asarray(fixed) - 1] = 1
    y = np.empty(m.n_dofs)
    stats = np.zeros(8)
    x = np.ascontiguousarray(x, dtype=np.float64)
    rc = L.hostcheck_matvec(m.elem_type, m.n_nodes, m.n_elems, coords.ctypes.data, conn.ctypes.data, fx.ctypes.data, EP, window, affine, E, nu,
                            x.ctypes.data, y.ctypes.data, project, stats.ctypes.data, n_owned)
    assert rc == 0, L.hostcheck_error().decode()
    return y, stats


@pytest.mark.parametrize("EP,window", [(256, 48), (128, 0), (256, 0)])
def test_tet10_affine_tables(hostcheck, oracle, jf, EP, window):
    m = jf.mesh.tet10_kuhn(9, 5, 4, 2.0, 1.0, 1.0)          # 1080 elements: several patches, one partial
    fixed = jf.mesh.clamp_dofs(m)
    u = jf.mesh.test_vector(m.n_dofs, fixed)
    y, st = run(hostcheck, m, u, fixed, EP=EP, window=window, project=1)
    ref = oracle.matfree(10, m.coords, m.conn, u, par=(210e9, 0.3), fixed_dofs=fixed)
    assert st[7] == m.n_elems                                 # every element classified affine
    assert relerr(y, ref) < 1e-12
    assert st[1] <= st[0] + 1e-9                              # lane assignment never increases the modelled conflicts


def test_tet10_mixed_curved_and_affine(hostcheck, oracle, jf):
    m = jf.mesh.tet10_kuhn(8, 4, 4, 2.0, 1.0, 1.0)
    rng = np.random.default_rng(0)
    sel = m.coords[:, 0] > 1.0
    m.coords[sel] += 2e-3 * rng.standard_normal((int(sel.sum()), 3))
    u = jf.mesh.test_vector(m.n_dofs)
    y, st = run(hostcheck, m, u)
    ref = oracle.matfree(10, m.coords, m.conn, u, par=(210e9, 0.3))
    assert 0 < st[7] < m.n_elems                              # both classes present -> interface nodes shared between the two sets
    assert relerr(y, ref) < 1e-12


def test_tet10_curved_general(hostcheck, oracle, jf):
    m = curved_tet10(jf.mesh)
    u = jf.mesh.test_vector(m.n_dofs)
    y, _ = run(hostcheck, m, u, EP=128)
    assert relerr(y, oracle.matfree(10, m.coords, m.conn, u, par=(210e9, 0.3))) < 1e-12


def test_hex8_and_tet4(hostcheck, oracle, jf):
    m = distorted_hex8(jf.mesh, n=9)
    u = jf.mesh.test_vector(m.n_dofs)
    y, st = run(hostcheck, m, u, EP=128)
    assert st[7] == 0                                         # distorted: no parallelepipeds
    assert relerr(y, oracle.matfree(8, m.coords, m.conn, u, par=(210e9, 0.3))) < 1e-12
    # parallelepiped elements (lattice sheared and stretched by an affine map) take the closed form; half of the mesh distorted
    m = jf.mesh.hex8_lattice(9, 7, 8, 0.25)
    A = np.array([[1.0, 0.3, 0.1], [0.0, 0.8, 0.2], [0.05, 0.0, 1.3]])
    m.coords = m.coords @ A.T
    rng = np.random.default_rng(1)
    sel = m.coords[:, 2] > 1.2
    m.coords[sel] += 0.01 * rng.standard_normal((int(sel.sum()), 3))
    u = jf.mesh.test_vector(m.n_dofs)
    y, st = run(hostcheck, m, u, EP=128)
    assert 0 < st[7] < m.n_elems
    assert relerr(y, oracle.matfree(8, m.coords, m.conn, u, par=(210e9, 0.3))) < 1e-12
    m = jf.mesh.tet4_kuhn(5, 4, 3, 1.0)
    u = jf.mesh.test_vector(m.n_dofs)
    y, _ = run(hostcheck, m, u, EP=128)
    assert relerr(y, oracle.matfree(4, m.coords, m.conn, u, par=(210e9, 0.3))) < 1e-12


def test_unstructured_fixture_and_orphan_nodes(hostcheck, oracle, jf):
    d = np.load(os.path.join(HERE, "golden", "tet10_fixture.npz"))
    coords, conn = d["coords"], d["conn"]
    # append two nodes no element references: y must be exactly zero there
    coords = np.vstack([coords, [[9.0, 9.0, 9.0], [8.0, 8.0, 8.0]]])
    m = jf.mesh.Mesh(10, coords, conn.astype(np.int32))
    u = jf.mesh.test_vector(m.n_dofs)
    y, _ = run(hostcheck, m, u, EP=128)
    ref = oracle.matfree(10, m.coords, m.conn, u, par=(210e9, 0.3))
    assert relerr(y, ref) < 1e-12
    assert np.all(y[-6:] == 0.0)


def test_lane_assignment_reduces_modelled_conflicts(hostcheck, jf):
    m = jf.mesh.tet10_kuhn(16, 8, 8, 2.0, 1.0, 1.0)
    u = jf.mesh.test_vector(m.n_dofs)
    _, st = run(hostcheck, m, u)
    before, after, ideal = st[0], st[1], st[2]
    assert ideal <= after < before
    print(f"modelled 64-bit shared-memory wavefronts per patch and component: {before:.0f} -> {after:.0f} (ideal {ideal:.0f})")


def test_ghost_aware_patch_order_and_landing_buffer(hostcheck, oracle, jf):
    """Partitioned mesh (real partition of a 2-rank run, local numbering): patches that read ghost nodes (ids >= n_owned)
    are ordered last and flagged; ghost values are taken from the landing buffer in ghost order (x's ghost entries are
    poisoned in the replay); owned rows equal the global operator."""
    m = jf.mesh.tet10_kuhn(6, 5, 12, 1.0, 1.0, 2.0)
    u = jf.mesh.test_vector(m.n_dofs)
    yref = oracle.matfree(10, m.coords, m.conn, u, par=(210e9, 0.3)).reshape(-1, 3)
    for world in (2, 3, 4):              # 3, 4: middle ranks with ghost nodes below AND above the owned range (two neighbours)
        for rank in range(world):
            p = jf.mesh.partition_mesh(m, world, rank)
            ml = jf.mesh.Mesh(10, m.coords[p.local_nodes - 1], p.conn_local)
            ul = u.reshape(-1, 3)[p.local_nodes - 1].ravel()
            y, _ = run(hostcheck, ml, ul, n_owned=p.n_owned)
            own = y.reshape(-1, 3)[:p.n_owned]
            assert relerr(own, yref[p.local_nodes[:p.n_owned] - 1]) < 1e-12
            if 0 < rank < world - 1:
                assert len(p.recv) == 2 and len(p.send) == 2


def test_tiny_and_empty_meshes(hostcheck, oracle, jf):
    """Edge cases: one element (a single, partial patch), two elements sharing a face, and a mesh without elements
    (every node is an orphan: y = 0)."""
    for et, m in ((10, jf.mesh.tet10_kuhn(1, 1, 1)), (8, jf.mesh.hex8_lattice(2, 2, 2, 0.5)), (4, jf.mesh.tet4_kuhn(1, 1, 1))):
        for ne in (1, min(2, m.n_elems)):
            sub = jf.mesh.Mesh(et, m.coords, m.conn[:ne])
            u = jf.mesh.test_vector(sub.n_dofs)
            y, st = run(hostcheck, sub, u, EP=128)
            assert st[3] == 1                                     # one patch
            assert relerr(y, oracle.matfree(et, sub.coords, sub.conn, u, par=(210e9, 0.3))) < 1e-12
    empty = jf.mesh.Mesh(10, jf.mesh.tet10_kuhn(1, 1, 1).coords, np.zeros((0, 10), dtype=np.int32))
    y, st = run(hostcheck, empty, np.ones(empty.n_dofs))
    assert st[3] == 0 and np.all(y == 0.0)
