"""CPU-side validation of the warp-per-element assembly arithmetic (csrc/asm_elem.cuh, used by elem_warp_kernel in
csrc/assemble.cu): tests/hostcheck replays one element "lane by lane" on the host and must reproduce the oracle's
element matrices Km (+Kg) for every element type and material.  Tolerance 1e-12 relative (north_star: stiffness entries).
The CUDA kernel itself is covered by the -m gpu tests (element_matrices / assemble_csr parity)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import relerr

HERE = os.path.dirname(os.path.abspath(__file__))
LE = (210e9, 0.3)
NH = (1.0e6, 0.3)
PP = (200e9, 0.3, 100e6, 10e9)


@pytest.fixture(scope="module")
def hostcheck():
    subprocess.check_call(["make", "-C", os.path.join(HERE, "hostcheck"), "-s"], stderr=subprocess.DEVNULL)
    L = C.CDLL(os.path.join(HERE, "hostcheck", "libhostcheck.so"))
    L.hostcheck_element_columns.restype = C.c_int
    L.hostcheck_element_columns.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    return L


def columns(L, et, X, u=None, kind=0, par=LE, geo=0, state_old=None):
    nd = 3 * et
    X = np.ascontiguousarray(X, dtype=np.float64)
    uu = np.zeros_like(X) if u is None else np.ascontiguousarray(u, dtype=np.float64)
    p4 = np.zeros(4)
    p4[:len(par)] = par
    st = None
    if state_old is not None:                      # oracle layout (ngp, 13) -> device layout SoA [13][ngp]
        st = np.ascontiguousarray(np.asarray(state_old, dtype=np.float64).T)
    Ke = np.zeros((nd, nd))
    rc = L.hostcheck_element_columns(et, X.ctypes.data, uu.ctypes.data, kind, p4.ctypes.data, geo, st.ctypes.data if st is not None else None,
                                     Ke.ctypes.data)
    return Ke.T.copy(), rc                          # column-major buffer -> [row, col]


def element_nodes(jf, et, rng, curved=True):
    if et == 10:
        m = jf.mesh.tet10_kuhn(1, 1, 1, 0.7, 1.1, 0.9)
    elif et == 8:
        m = jf.mesh.hex8_lattice(2, 2, 2, 0.8)
    else:
        m = jf.mesh.tet4_kuhn(1, 1, 1, 1.3)
    e = int(rng.integers(m.n_elems))
    X = m.coords[m.conn[e] - 1].copy()
    if curved:
        X += 0.03 * rng.standard_normal(X.shape)
    return X


@pytest.mark.parametrize("et", [10, 8, 4])
def test_linear_elastic_columns(hostcheck, oracle, jf, et):
    rng = np.random.default_rng(et)
    for curved in (False, True):
        X = element_nodes(jf, et, rng, curved)
        Km, _, _, _ = oracle.element(et, X, par=LE)
        Ke, rc = columns(hostcheck, et, X)
        assert rc == 0
        assert relerr(Ke, Km) < 1e-12
        assert relerr(Ke, Ke.T) < 1e-12


@pytest.mark.parametrize("et", [10, 8, 4])
def test_neo_hookean_and_stvk_tangent_columns(hostcheck, oracle, jf, et):
    rng = np.random.default_rng(10 + et)
    X = element_nodes(jf, et, rng)
    u = 0.02 * rng.standard_normal(X.shape)
    Km, Kg, _, _ = oracle.element(et, X, u, kind=1, par=NH, finite_strain=True, geometric=True)
    Ke, rc = columns(hostcheck, et, X, u, kind=1, par=NH)
    assert rc == 0 and relerr(Ke, Km + Kg) < 1e-11
    # St. Venant-Kirchhoff (the classic finite_strain path), with and without the geometric stiffness
    for geo in (0, 1):
        Km, Kg, _, _ = oracle.element(et, X, u, kind=0, par=LE, finite_strain=True, geometric=bool(geo))
        Ke, rc = columns(hostcheck, et, X, u, kind=3, par=LE, geo=geo)
        assert rc == 0 and relerr(Ke, Km + (Kg if geo else 0.0)) < 1e-12


def test_invalid_deformation_is_flagged(hostcheck, jf):
    rng = np.random.default_rng(5)
    X = element_nodes(jf, 10, rng, curved=False)
    u = -2.5 * X                                    # F = -1.5 I: det(C) > 0 but the oracle's J check is on C; use a collapse instead
    u = -1.0 * X                                    # F = 0: det(C) = 0 -> flagged
    _, rc = columns(hostcheck, 10, X, u, kind=1, par=NH)
    assert rc == 1


def test_plastic_tangent_columns(hostcheck, oracle, jf):
    """Consistent elastoplastic tangent at yielded Gauss points: strain large enough that every point is plastic, old state
    non-zero (taken from a first oracle step)."""
    rng = np.random.default_rng(3)
    X = element_nodes(jf, 10, rng)
    u1 = 2e-3 * rng.standard_normal(X.shape)
    _, _, _, s1 = oracle.element(10, X, u1, kind=2, par=PP)
    assert (s1[:, 12] > 0).all()                    # plastic multiplier accumulated at all four points
    u2 = u1 + 1e-3 * rng.standard_normal(X.shape)
    Km, _, _, _ = oracle.element(10, X, u2, kind=2, par=PP, state_old=s1)
    Ke, rc = columns(hostcheck, 10, X, u2, kind=2, par=PP, state_old=s1)
    assert rc == 0 and relerr(Ke, Km) < 1e-11
    # elastic step from a virgin state reproduces the linear-elastic matrix
    Ke0, _ = columns(hostcheck, 10, X, 1e-9 * u1, kind=2, par=PP, state_old=np.zeros((4, 13)))
    Kle, _, _, _ = oracle.element(10, X, par=PP[:2])
    assert relerr(Ke0, Kle) < 1e-12


def _pattern(L, m):
    L.hostcheck_pattern.restype = C.c_longlong
    L.hostcheck_pattern.argtypes = [C.c_int, C.c_longlong, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    conn = np.ascontiguousarray(m.conn - 1, dtype=np.int32)
    rowptr = np.zeros(3 * m.n_nodes + 1, dtype=np.int64)
    times = np.zeros(2)
    nnz = L.hostcheck_pattern(m.elem_type, m.n_nodes, m.n_elems, conn.ctypes.data, rowptr.ctypes.data, None, None, times.ctypes.data)
    assert nnz >= 0
    colind = np.zeros(nnz, dtype=np.int32)
    colour = np.full(m.n_elems, -1, dtype=np.int32)
    assert L.hostcheck_pattern(m.elem_type, m.n_nodes, m.n_elems, conn.ctypes.data, rowptr.ctypes.data, colind.ctypes.data, colour.ctypes.data,
                               times.ctypes.data) == nnz
    return rowptr, colind, colour, times


def test_pattern_and_colouring_host_side(hostcheck, oracle, jf):
    """The host part of the assembled path (node adjacency -> CSR pattern, greedy colouring) is bit-identical to the
    oracle's restatement of sparse(I, J, V) (src/sparse/sparse.jl:121-132) and of the greedy colouring
    (src/preprocess.jl:331-398), on structured, unstructured and orphan-node meshes."""
    d = np.load(os.path.join(HERE, "golden", "tet10_fixture.npz"))
    fixture = jf.mesh.Mesh(10, np.vstack([d["coords"], [[9.0, 9.0, 9.0]]]), d["conn"].astype(np.int32))   # + one orphan node
    for m in (jf.mesh.tet10_kuhn(5, 4, 3), jf.mesh.hex8_lattice(6, 5, 4, 0.2), jf.mesh.tet4_kuhn(4, 3, 5), fixture, jf.mesh.tet10_kuhn(1, 1, 1)):
        rowptr, colind, colour, _ = _pattern(hostcheck, m)
        rp, ci = oracle.csr_pattern(m.elem_type, m.n_nodes, m.conn)
        assert np.array_equal(rowptr, rp) and np.array_equal(colind, ci)
        col, n = oracle.colouring(m.elem_type, m.n_nodes, m.conn)
        assert np.array_equal(colour, col) and colour.max() + 1 == n
        # elements of one colour share no node
        for c in range(n):
            nodes = m.conn[colour == c].ravel()
            assert np.unique(nodes).size == nodes.size
