"""ctypes wrapper of the CPU oracle (oracle/jfem_oracle.c).

TEST INFRASTRUCTURE ONLY -- may be imported from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs, never from the product package.
Takes the reference's 1-based node ids and converts once.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libjfem_oracle.so")

MAT_LE, MAT_NH, MAT_PP = 0, 1, 2
NSTATE = 13
NGP = {4: 1, 8: 8, 10: 4}


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "jfem_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


_lib = None
_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_lp = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
_bp = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        L = _lib
        L.orc_num_threads.restype = C.c_int
        L.orc_set_num_threads.argtypes = [C.c_int]
        L.orc_shape_N.argtypes = [C.c_int, _dp, _dp]
        L.orc_shape_dN.argtypes = [C.c_int, _dp, _dp]
        L.orc_quadrature.argtypes = [C.c_int, _dp, _dp]
        L.orc_quadrature.restype = C.c_int
        L.orc_le_stress.argtypes = [C.c_double, C.c_double, _dp, _dp, _dp]
        L.orc_nh_energy.argtypes = [C.c_double, C.c_double, _dp]
        L.orc_nh_energy.restype = C.c_double
        L.orc_nh_stress.argtypes = [C.c_double, C.c_double, _dp, _dp, _dp]
        L.orc_nh_stress.restype = C.c_int
        L.orc_pp_stress.argtypes = [_dp, _dp, C.c_void_p, _dp, _dp, C.c_void_p]
        L.orc_pp_stress.restype = C.c_int
        L.orc_element.argtypes = [C.c_int, _dp, _dp, C.c_int, _dp, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_element.restype = C.c_int
        L.orc_element_block_form.argtypes = [C.c_int, _dp, C.c_double, C.c_double, _dp]
        L.orc_csr_pattern.argtypes = [C.c_int, C.c_int64, C.c_int64, _ip, _lp, C.c_void_p]
        L.orc_assemble_csr.argtypes = [C.c_int, C.c_int64, C.c_int64, _dp, _ip, C.c_void_p, C.c_int, _dp, C.c_int, C.c_int,
                                       C.c_void_p, C.c_void_p, _lp, _ip, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_assemble_csr.restype = C.c_int
        L.orc_spmv.argtypes = [C.c_int64, _lp, _ip, _dp, _dp, _dp]
        L.orc_matfree.argtypes = [C.c_int, C.c_int64, C.c_int64, _dp, _ip, _dp, C.c_int, _dp, C.c_int, C.c_void_p, C.c_void_p, _dp]
        L.orc_matfree.restype = C.c_int
        L.orc_cg_csr.argtypes = [C.c_int64, _lp, _ip, _dp, C.c_void_p, _dp, _dp, C.c_double, C.c_int, C.c_int, C.POINTER(C.c_double)]
        L.orc_cg_csr.restype = C.c_int
        L.orc_colouring.argtypes = [C.c_int, C.c_int64, C.c_int64, _ip, _ip]
        L.orc_colouring.restype = C.c_int
        L.orc_body_load.argtypes = [C.c_int, C.c_int64, C.c_int64, _dp, _ip, _dp, _dp]
        L.orc_surface_load.argtypes = [C.c_int, C.c_int64, _dp, _ip, C.c_void_p, C.c_void_p, _dp]
    return _lib


def _par(par):
    p = np.zeros(4)
    p[: len(par)] = par
    return p


def _vp(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def num_threads() -> int:
    return lib().orc_num_threads()


def set_num_threads(n: int) -> int:
    """OpenMP threads of the oracle (a process started by torchrun inherits OMP_NUM_THREADS=1)."""
    lib().orc_set_num_threads(int(n))
    return num_threads()


def shape_N(et, xi):
    N = np.zeros(et)
    lib().orc_shape_N(et, np.ascontiguousarray(xi, dtype=np.float64), N)
    return N


def shape_dN(et, xi):
    dN = np.zeros((et, 3))
    lib().orc_shape_dN(et, np.ascontiguousarray(xi, dtype=np.float64), dN)
    return dN


def quadrature(et):
    w, xi = np.zeros(8), np.zeros((8, 3))
    n = lib().orc_quadrature(et, w, xi)
    return w[:n].copy(), xi[:n].copy()


def le_stress(E, nu, eps):
    s, D = np.zeros(6), np.zeros((6, 6))
    lib().orc_le_stress(E, nu, np.ascontiguousarray(eps, dtype=np.float64), s, D)
    return s, D


def lame(E, nu):
    return E * nu / ((1.0 + nu) * (1.0 - 2.0 * nu)), E / (2.0 * (1.0 + nu))


def nh_energy(mu, la, Cv):
    return lib().orc_nh_energy(mu, la, np.ascontiguousarray(Cv, dtype=np.float64))


def nh_stress(mu, la, Ev):
    S, D = np.zeros(6), np.zeros((6, 6))
    rc = lib().orc_nh_stress(mu, la, np.ascontiguousarray(Ev, dtype=np.float64), S, D)
    if rc:
        raise ValueError("Jacobian J = sqrt(det(C)) must be positive")  # neo_hookean.jl:137
    return S, D


def pp_stress(par, eps, state_old=None):
    s, D, st = np.zeros(6), np.zeros((6, 6)), np.zeros(NSTATE)
    so = None if state_old is None else np.ascontiguousarray(state_old, dtype=np.float64)
    plastic = lib().orc_pp_stress(_par(par), np.ascontiguousarray(eps, dtype=np.float64), _vp(so), s, D, _vp(st))
    return s, D, st, bool(plastic)


def element(et, X, u=None, kind=MAT_LE, par=(210e9, 0.3), finite_strain=False, geometric=False, state_old=None):
    """Returns (Km, Kg, fint, state_new); Km/Kg are (ndof, ndof) numpy arrays indexed [row, col]."""
    nd = 3 * et
    X = np.ascontiguousarray(X, dtype=np.float64)
    u = np.zeros_like(X) if u is None else np.ascontiguousarray(u, dtype=np.float64).reshape(et, 3)
    Km, Kg, f = np.zeros((nd, nd)), np.zeros((nd, nd)), np.zeros(nd)
    so = None if state_old is None else np.ascontiguousarray(state_old, dtype=np.float64)
    sn = np.zeros((NGP[et], NSTATE))
    rc = lib().orc_element(et, X, u, kind, _par(par), int(finite_strain), int(geometric), _vp(so), _vp(sn), _vp(Km), _vp(Kg), _vp(f))
    if rc:
        raise ValueError("invalid deformation (J <= 0)")
    return Km.T.copy(), Kg.T.copy(), f, sn   # C buffer is column-major -> transpose to [row, col]


def element_block_form(et, X, E, nu):
    nd = 3 * et
    Ke = np.zeros((nd, nd))
    lib().orc_element_block_form(et, np.ascontiguousarray(X, dtype=np.float64), E, nu, Ke)
    return Ke.T.copy()


def _conn0(conn):
    return np.ascontiguousarray(np.asarray(conn, dtype=np.int32) - 1)


def csr_pattern(et, n_nodes, conn):
    """(rowptr int64 0-based, colind int32 0-based) of the reference's sparse(K)."""
    c0 = _conn0(conn)
    rowptr = np.zeros(3 * n_nodes + 1, dtype=np.int64)
    lib().orc_csr_pattern(et, n_nodes, c0.shape[0], c0, rowptr, None)
    colind = np.zeros(int(rowptr[-1]), dtype=np.int32)
    lib().orc_csr_pattern(et, n_nodes, c0.shape[0], c0, rowptr, _vp(colind))
    return rowptr, colind


def assemble_csr(et, coords, conn, u=None, kind=MAT_LE, par=(210e9, 0.3), finite_strain=False, geometric=False,
                 state_old=None, pattern=None, symmetrise=False, want_state=False):
    coords = np.ascontiguousarray(coords, dtype=np.float64)
    c0 = _conn0(conn)
    nn = coords.shape[0]
    rowptr, colind = pattern if pattern is not None else csr_pattern(et, nn, conn)
    vals, f = np.zeros(int(rowptr[-1])), np.zeros(3 * nn)
    uu = None if u is None else np.ascontiguousarray(u, dtype=np.float64)
    so = None if state_old is None else np.ascontiguousarray(state_old, dtype=np.float64)
    sn = np.zeros((c0.shape[0], NGP[et], NSTATE)) if want_state else None
    rc = lib().orc_assemble_csr(et, nn, c0.shape[0], coords, c0, _vp(uu), kind, _par(par), int(finite_strain), int(geometric),
                                _vp(so), _vp(sn), rowptr, colind, _vp(vals), _vp(f), int(symmetrise))
    if rc:
        raise ValueError("invalid deformation (J <= 0)")
    return (rowptr, colind, vals, f, sn) if want_state else (rowptr, colind, vals, f)


def spmv(rowptr, colind, vals, x):
    y = np.zeros(rowptr.size - 1)
    lib().orc_spmv(rowptr.size - 1, rowptr, colind, vals, np.ascontiguousarray(x, dtype=np.float64), y)
    return y


def matfree(et, coords, conn, u, kind=MAT_LE, par=(210e9, 0.3), finite_strain=False, state_old=None, fixed_dofs=None):
    """f_int(u) (== K.u for small-strain LE).  fixed_dofs: 1-based dof ids zeroed on output."""
    coords = np.ascontiguousarray(coords, dtype=np.float64)
    c0 = _conn0(conn)
    nn = coords.shape[0]
    y = np.zeros(3 * nn)
    mask = None
    if fixed_dofs is not None:
        mask = np.zeros(3 * nn, dtype=np.uint8)
        mask[np.asarray(fixed_dofs, dtype=np.int64) - 1] = 1
    so = None if state_old is None else np.ascontiguousarray(state_old, dtype=np.float64)
    rc = lib().orc_matfree(et, nn, c0.shape[0], coords, c0, np.ascontiguousarray(u, dtype=np.float64), kind, _par(par),
                           int(finite_strain), _vp(so), _vp(mask), y)
    if rc:
        raise ValueError("invalid deformation (J <= 0)")
    return y


def cg_csr(rowptr, colind, vals, b, fixed_dofs=None, x0=None, tol=1e-6, relative=False, max_iter=1000):
    n = rowptr.size - 1
    mask = None
    if fixed_dofs is not None:
        mask = np.zeros(n, dtype=np.uint8)
        mask[np.asarray(fixed_dofs, dtype=np.int64) - 1] = 1
    x = np.zeros(n) if x0 is None else np.array(x0, dtype=np.float64)
    res = C.c_double(0.0)
    it = lib().orc_cg_csr(n, rowptr, colind, vals, _vp(mask), np.ascontiguousarray(b, dtype=np.float64), x, tol, int(relative),
                          max_iter, C.byref(res))
    return x, it, res.value


def colouring(et, n_nodes, conn):
    c0 = _conn0(conn)
    col = np.zeros(c0.shape[0], dtype=np.int32)
    n = lib().orc_colouring(et, n_nodes, c0.shape[0], c0, col)
    return col, n


def body_load(et, coords, conn, b):
    """Consistent body load f_ext (src/problems_elasticity.jl:412-426); b = 3 values or (n_elems, 3)."""
    coords = np.ascontiguousarray(coords, dtype=np.float64)
    c0 = _conn0(conn)
    bb = np.ascontiguousarray(np.broadcast_to(np.asarray(b, dtype=np.float64), (c0.shape[0], 3)))
    f = np.zeros(3 * coords.shape[0])
    lib().orc_body_load(et, coords.shape[0], c0.shape[0], coords, c0, bb, f)
    return f


def surface_load(face_type, coords, faces, traction=None, pressure=None, n_dofs=None):
    """Consistent surface traction / pressure load vector (src/problems_elasticity.jl:454-502); faces: (n_faces, nn) 1-based."""
    coords = np.ascontiguousarray(coords, dtype=np.float64)
    f0 = _conn0(np.asarray(faces).reshape(-1, face_type))
    nf = f0.shape[0]
    t = None if traction is None else np.ascontiguousarray(np.broadcast_to(np.asarray(traction, dtype=np.float64), (nf, 3)))
    p = None if pressure is None else np.ascontiguousarray(np.broadcast_to(np.asarray(pressure, dtype=np.float64), (nf,)))
    f = np.zeros(3 * coords.shape[0] if n_dofs is None else n_dofs)
    lib().orc_surface_load(face_type, nf, coords, f0, _vp(t), _vp(p), f)
    return f


def quadrature_mass(et):
    """Rule of the least-squares recovery: the next rule of the reference's integration_rule_mapping
    (src/elements/integrate.jl:11-31) that integrates N N' exactly -- Tet4: GLTET4, Tet10: GLTET15
    (src/quadrature/gltet.jl:44-64), Hex8: GLHEX8.  (The default rules make the Tet4 / Tet10 mass matrix singular.)"""
    if et == 8:
        return quadrature(8)
    if et == 4:
        return quadrature(10)
    s15 = np.sqrt(15.0)
    a, b1, b2 = 0.25, (7.0 + s15) / 34.0, (7.0 - s15) / 34.0
    c1, c2, d, f = (13.0 - 3.0 * s15) / 34.0, (13.0 + 3.0 * s15) / 34.0, (5.0 - s15) / 20.0, (5.0 + s15) / 20.0
    w1, w2, w3, w4 = 8.0 / 405.0, (2665.0 - 14.0 * s15) / 226800.0, (2665.0 + 14.0 * s15) / 226800.0, 5.0 / 567.0
    pts = np.array([(a, a, a), (b1, b1, b1), (b1, b1, c1), (b1, c1, b1), (c1, b1, b1), (b2, b2, b2), (b2, b2, c2), (b2, c2, b2),
                    (c2, b2, b2), (d, d, f), (d, f, d), (f, d, d), (d, f, f), (f, d, f), (f, f, d)])
    return np.array([w1] + [w2] * 4 + [w3] * 4 + [w4] * 6), pts


def lsq_recover(et, coords, conn, u, field="stress", par=(210e9, 0.3), default_rule=False):
    """Least-squares nodal fit of the Gauss-point strain / stress (lsq_fit, src/problems_elasticity.jl:547-594):
    A = sum w detJ N N', b_i = sum w detJ f_i N, A <- (A+A')/2, x = A \\ b (direct solve; the reference uses ldlt).
    f = strain vector [e11,e22,e33,e12,e23,e13] with eps = (grad u' + grad u)/2 (:520-524,540-545), or the
    linear-elastic stress la tr(eps) I + 2 mu eps of it (:527-537).  Returns (n_nodes, 6)."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    coords = np.ascontiguousarray(coords, dtype=np.float64)
    c0 = _conn0(conn).astype(np.int64)
    nn, ne = coords.shape[0], c0.shape[0]
    X = coords[c0]                                      # (ne, nnpe, 3)
    U = np.asarray(u, dtype=np.float64).reshape(-1, 3)[c0]
    w, xi = quadrature(et) if default_rule else quadrature_mass(et)
    la, mu = lame(*par[:2])
    Ae = np.zeros((ne, et, et))
    be = np.zeros((ne, et, 6))
    for wg, x in zip(w, xi):
        N, dN = shape_N(et, x), shape_dN(et, x)          # (nnpe,), (nnpe, 3)
        J = np.einsum("ia,eib->eab", dN, X)              # J[a,b] = sum dN_i[a] X_i[b]
        detJ = np.linalg.det(J)
        G = np.einsum("eab,ib->eia", np.linalg.inv(J), dN)   # grad N_i = inv(J) dN_i
        gu = np.einsum("eic,eid->ecd", U, G)             # grad u = sum u_k (x) grad N_k
        eps = 0.5 * (gu + np.swapaxes(gu, 1, 2))
        if field == "stress":
            t = la * np.trace(eps, axis1=1, axis2=2)
            eps = 2 * mu * eps + t[:, None, None] * np.eye(3)[None]
        fv = np.stack([eps[:, 0, 0], eps[:, 1, 1], eps[:, 2, 2], eps[:, 0, 1], eps[:, 1, 2], eps[:, 0, 2]], axis=1)
        wd = wg * detJ
        Ae += wd[:, None, None] * (N[:, None] * N[None, :])[None]
        be += wd[:, None, None] * N[None, :, None] * fv[:, None, :]
    rows = np.repeat(c0, et, axis=1).ravel()
    cols = np.tile(c0, (1, et)).ravel()
    A = sp.coo_matrix((Ae.ravel(), (rows, cols)), shape=(nn, nn)).tocsr()
    A = 0.5 * (A + A.T)
    b = np.zeros((nn, 6))
    np.add.at(b, c0.ravel(), be.reshape(-1, 6))
    nz = np.nonzero(np.asarray(abs(A).sum(axis=1)).ravel() > 0)[0]        # get_nonzero_rows
    x = np.zeros((nn, 6))
    x[nz] = spla.splu(A[nz][:, nz].tocsc()).solve(b[nz])
    return x
