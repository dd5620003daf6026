"""ctypes binding of libjfem_b200.so (the C ABI of include/jfem_b200.h).

There is NO CPU fallback: if the shared library is missing or no B200 is visible, calls raise.
numpy arrays are passed as host pointers; torch CUDA tensors (float64, contiguous) as device pointers.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libjfem_b200.so")

TET4, HEX8, TET10 = 4, 8, 10
MAT_LINEAR_ELASTIC, MAT_NEO_HOOKEAN, MAT_PERFECT_PLASTICITY, MAT_STVK = 0, 1, 2, 3
FIELD_STRAIN, FIELD_STRESS = 0, 1
TRI3, QUAD4, TRI6 = 3, 4, 6
PROJECT, TANGENT, USE_CSR, JACOBI = 1, 2, 4, 8
NSTATE = 13
NGP = {4: 1, 8: 8, 10: 4}

EXPORTS = [
    "jfem_abi_version", "jfem_last_error", "jfem_device_count", "jfem_create", "jfem_destroy", "jfem_set_option",
    "jfem_set_material", "jfem_set_dirichlet", "jfem_get_info", "jfem_set_stream", "jfem_synchronize", "jfem_matvec",
    "jfem_internal_force", "jfem_set_linearization", "jfem_commit_state", "jfem_get_state", "jfem_set_state",
    "jfem_element_matrices", "jfem_csr_size", "jfem_csr_pattern", "jfem_assemble_csr", "jfem_spmv", "jfem_cg",
    "jfem_newton_krylov", "jfem_comm_unique_id", "jfem_comm_init", "jfem_comm_set_halo", "jfem_comm_p2p_export",
    "jfem_comm_p2p_import", "jfem_comm_p2p_seq", "jfem_comm_destroy", "jfem_body_load", "jfem_surface_load", "jfem_reactions",
    "jfem_nodal_recover", "jfem_csr_penalty_bc",
]


class JfemError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libjfem_b200 error {code}: {msg}")
        self.code = code


class DomainError(JfemError):
    """Invalid deformation, J <= 0 (the reference throws DomainError, src/materials/neo_hookean.jl:137)."""


class Info(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("device", C.c_int32), ("elem_type", C.c_int32), ("n_ranks", C.c_int32),
                ("n_nodes", C.c_int64), ("n_elems", C.c_int64), ("n_dofs", C.c_int64), ("n_fixed", C.c_int64),
                ("n_patches", C.c_int64), ("n_interface_nodes", C.c_int64), ("n_affine_elems", C.c_int64),
                ("patch_elems", C.c_int64), ("patch_max_nodes", C.c_int64), ("device_bytes", C.c_int64),
                ("matvec_launches", C.c_int64), ("total_launches", C.c_int64), ("smem_bytes", C.c_int64),
                ("blocks_per_sm", C.c_int64), ("setup_seconds", C.c_double)]


_lib = None


def lib():
    """Load the shared library (built in-tree by `make -C juliafem.jl_b200/csrc` / __graft_entry__.build())."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FileNotFoundError(f"{LIB_PATH} not found: build it with `make -C {os.path.join(_HERE, 'csrc')}` "
                                    "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        vp, i32, i64, dbl = C.c_void_p, C.c_int, C.c_int64, C.c_double
        L.jfem_abi_version.restype = i32
        L.jfem_last_error.restype = C.c_char_p
        L.jfem_device_count.argtypes = [C.POINTER(i32)]
        L.jfem_create.argtypes = [C.POINTER(vp), i32, i32, i64, i64, vp, vp, i32]
        L.jfem_destroy.argtypes = [vp]
        L.jfem_set_option.argtypes = [vp, C.c_char_p, dbl]
        L.jfem_set_material.argtypes = [vp, i32, vp, i32, i32]
        L.jfem_set_dirichlet.argtypes = [vp, vp, vp, i64]
        L.jfem_get_info.argtypes = [vp, C.POINTER(Info)]
        L.jfem_set_stream.argtypes = [vp, vp]
        L.jfem_synchronize.argtypes = [vp]
        L.jfem_matvec.argtypes = [vp, vp, vp, i32, i32]
        L.jfem_internal_force.argtypes = [vp, vp, vp, i32, i32]
        L.jfem_set_linearization.argtypes = [vp, vp, i32]
        L.jfem_commit_state.argtypes = [vp]
        L.jfem_get_state.argtypes = [vp, vp, i32]
        L.jfem_set_state.argtypes = [vp, vp]
        L.jfem_element_matrices.argtypes = [vp, vp, i64, i64, vp, vp]
        L.jfem_csr_size.argtypes = [vp, C.POINTER(i64), C.POINTER(i64)]
        L.jfem_csr_pattern.argtypes = [vp, vp, vp]
        L.jfem_assemble_csr.argtypes = [vp, vp, vp, vp, i32, i32]
        L.jfem_spmv.argtypes = [vp, vp, vp, i32, i32]
        L.jfem_cg.argtypes = [vp, vp, vp, dbl, i32, i32, i32, C.POINTER(i32), C.POINTER(dbl), i32]
        L.jfem_newton_krylov.argtypes = [vp, vp, vp, dbl, i32, i32, dbl, dbl, i32, C.POINTER(i32), C.POINTER(i32),
                                         C.POINTER(dbl), vp, i32, i32]
        L.jfem_comm_unique_id.argtypes = [C.c_char_p]
        L.jfem_comm_init.argtypes = [vp, i32, i32, C.c_char_p, i64]
        L.jfem_comm_set_halo.argtypes = [vp, i32, vp, vp, vp, vp, vp]
        L.jfem_comm_p2p_export.argtypes = [vp, C.c_char_p]
        L.jfem_comm_p2p_import.argtypes = [vp, C.c_char_p, vp, vp]
        L.jfem_comm_p2p_seq.argtypes = [vp, i64, C.POINTER(i64)]
        L.jfem_comm_destroy.argtypes = [vp]
        L.jfem_body_load.argtypes = [vp, vp, i32, i32, vp, i32]
        L.jfem_surface_load.argtypes = [vp, i32, i64, vp, vp, vp, i32, vp, i32]
        L.jfem_reactions.argtypes = [vp, vp, vp, vp, i32]
        L.jfem_nodal_recover.argtypes = [vp, vp, i32, vp, i32]
        L.jfem_csr_penalty_bc.argtypes = [vp, dbl, vp, C.POINTER(dbl), i32]
        for name in EXPORTS:
            if name != "jfem_last_error":
                getattr(L, name).restype = i32
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        msg = lib().jfem_last_error().decode(errors="replace")
        raise (DomainError if rc == 5 else JfemError)(rc, msg)


def device_count() -> int:
    n = C.c_int(0)
    lib().jfem_device_count(C.byref(n))
    return n.value


def _is_torch(x):
    return type(x).__module__.startswith("torch")


def _ptr(x, n=None, dtype=np.float64):
    """(pointer, on_device, keepalive)"""
    if x is None:
        return None, 0, None
    if _is_torch(x):
        import torch
        want = {np.float64: torch.float64}[dtype]
        if x.dtype != want or not x.is_contiguous():
            raise TypeError("device vectors must be contiguous float64 tensors")
        if n is not None and x.numel() != n:
            raise ValueError(f"vector length {x.numel()} != {n}")
        return C.c_void_p(x.data_ptr()), (1 if x.is_cuda else 0), x
    a = np.ascontiguousarray(x, dtype=dtype)
    if n is not None and a.size != n:
        raise ValueError(f"vector length {a.size} != {n}")
    return a.ctypes.data_as(C.c_void_p), 0, a


class Handle:
    """Owner of one jfem_handle (one mesh partition on one GPU)."""

    def __init__(self, elem_type, coords, conn, device=0, index_base=1):
        coords = np.ascontiguousarray(coords, dtype=np.float64).reshape(-1, 3)
        conn = np.ascontiguousarray(conn, dtype=np.int32).reshape(-1, elem_type)
        self.elem_type, self.n_nodes, self.n_elems = int(elem_type), coords.shape[0], conn.shape[0]
        self.n_dofs = 3 * self.n_nodes
        self.index_base = index_base
        self._h = C.c_void_p()
        check(lib().jfem_create(C.byref(self._h), device, elem_type, self.n_nodes, self.n_elems,
                                coords.ctypes.data_as(C.c_void_p), conn.ctypes.data_as(C.c_void_p), index_base))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().jfem_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def set_option(self, key, value):
        check(lib().jfem_set_option(self._h, key.encode(), float(value)))

    def set_material(self, kind, params):
        """params: (n_params,) homogeneous, or (n_elems, n_params) per element (row e = parameters of element e)."""
        p = np.ascontiguousarray(params, dtype=np.float64)
        if p.ndim == 2:
            assert p.shape[0] == self.n_elems
            check(lib().jfem_set_material(self._h, kind, p.ctypes.data_as(C.c_void_p), p.shape[1], 1))
        else:
            check(lib().jfem_set_material(self._h, kind, p.ctypes.data_as(C.c_void_p), p.size, 0))

    def set_dirichlet(self, dofs, values=None):
        d = np.ascontiguousarray(dofs, dtype=np.int64)
        v = None if values is None else np.ascontiguousarray(values, dtype=np.float64)
        check(lib().jfem_set_dirichlet(self._h, d.ctypes.data_as(C.c_void_p), None if v is None else v.ctypes.data_as(C.c_void_p), d.size))

    def info(self) -> Info:
        i = Info()
        check(lib().jfem_get_info(self._h, C.byref(i)))
        return i

    def set_stream(self, cuda_stream_ptr):
        check(lib().jfem_set_stream(self._h, C.c_void_p(cuda_stream_ptr)))

    def synchronize(self):
        check(lib().jfem_synchronize(self._h))

    def _inout(self, fn, x, y, flags):
        px, dx, kx = _ptr(x, self.n_dofs)
        if y is None:
            if dx:
                import torch
                y = torch.empty_like(x)
            else:
                y = np.empty(self.n_dofs)
        py, dy, ky = _ptr(y, self.n_dofs)
        if dx != dy:
            raise TypeError("x and y must both be host arrays or both device tensors")
        check(fn(self._h, px, py, flags, dx))
        return y

    def matvec(self, x, y=None, flags=0):
        return self._inout(lib().jfem_matvec, x, y, flags)

    def spmv(self, x, y=None, flags=0):
        return self._inout(lib().jfem_spmv, x, y, flags)

    def internal_force(self, u, f=None, flags=0):
        return self._inout(lib().jfem_internal_force, u, f, flags)

    def set_linearization(self, u):
        p, d, k = _ptr(u, self.n_dofs)
        check(lib().jfem_set_linearization(self._h, p, d))

    def commit_state(self):
        check(lib().jfem_commit_state(self._h))

    def get_state(self, committed=True):
        st = np.zeros((self.n_elems, NGP[self.elem_type], NSTATE))
        check(lib().jfem_get_state(self._h, st.ctypes.data_as(C.c_void_p), int(committed)))
        return st

    def set_state(self, st):
        st = np.ascontiguousarray(st, dtype=np.float64)
        assert st.size == self.n_elems * NGP[self.elem_type] * NSTATE
        check(lib().jfem_set_state(self._h, st.ctypes.data_as(C.c_void_p)))

    def element_matrices(self, u=None, e0=0, ne=None, want_K=True, want_f=True):
        """Ke[e] indexed [row, col] and fe[e] for caller elements e0..e0+ne (0-based range)."""
        ne = self.n_elems - e0 if ne is None else ne
        nd = 3 * self.elem_type
        K = np.zeros((ne, nd, nd)) if want_K else None
        f = np.zeros((ne, nd)) if want_f else None
        pu, _, ku = _ptr(u, self.n_dofs)
        check(lib().jfem_element_matrices(self._h, pu, e0, ne, None if K is None else K.ctypes.data_as(C.c_void_p),
                                          None if f is None else f.ctypes.data_as(C.c_void_p)))
        if K is not None:
            K = np.ascontiguousarray(np.transpose(K, (0, 2, 1)))   # column-major per element -> [row, col]
        return K, f

    def csr_size(self):
        """(n_rows, nnz) of the assembled operator; builds the pattern on the device without copying it to the host."""
        n, nnz = C.c_int64(0), C.c_int64(0)
        check(lib().jfem_csr_size(self._h, C.byref(n), C.byref(nnz)))
        return int(n.value), int(nnz.value)

    def csr_pattern(self):
        n, nnz = C.c_int64(0), C.c_int64(0)
        check(lib().jfem_csr_size(self._h, C.byref(n), C.byref(nnz)))
        rowptr = np.zeros(n.value + 1, dtype=np.int64)
        colind = np.zeros(nnz.value, dtype=np.int32)
        check(lib().jfem_csr_pattern(self._h, rowptr.ctypes.data_as(C.c_void_p), colind.ctypes.data_as(C.c_void_p)))
        return rowptr, colind

    def assemble_csr(self, u=None, symmetrise=False, want_vals=True, want_f=False):
        n, nnz = C.c_int64(0), C.c_int64(0)
        check(lib().jfem_csr_size(self._h, C.byref(n), C.byref(nnz)))
        pu, du, ku = _ptr(u, self.n_dofs)
        if du:
            check(lib().jfem_assemble_csr(self._h, pu, None, None, int(symmetrise), 1))
            return None, None
        vals = np.zeros(nnz.value) if want_vals else None
        f = np.zeros(self.n_dofs) if want_f else None
        check(lib().jfem_assemble_csr(self._h, pu, None if vals is None else vals.ctypes.data_as(C.c_void_p),
                                      None if f is None else f.ctypes.data_as(C.c_void_p), int(symmetrise), 0))
        return vals, f

    # ---- loads, reactions, post-processing (host arrays)
    def body_load(self, b, f=None):
        """f (+)= consistent body load; b = 3 values or (n_elems, 3).  f given -> accumulated into a copy of it."""
        bb = np.ascontiguousarray(b, dtype=np.float64)
        per = int(bb.ndim == 2)
        if per:
            assert bb.shape == (self.n_elems, 3)
        out = np.zeros(self.n_dofs) if f is None else np.array(f, dtype=np.float64)
        check(lib().jfem_body_load(self._h, bb.ctypes.data_as(C.c_void_p), per, int(f is not None), out.ctypes.data_as(C.c_void_p), 0))
        return out

    def surface_load(self, face_type, faces, traction=None, pressure=None, f=None):
        """Consistent surface traction (3 values or (n_faces, 3)) / pressure (scalar or (n_faces,)) on faces given by
        volume-mesh node ids (n_faces, face_type)."""
        fc = np.ascontiguousarray(faces, dtype=np.int32).reshape(-1, face_type)
        nf = fc.shape[0]
        t = None if traction is None else np.ascontiguousarray(np.broadcast_to(np.asarray(traction, dtype=np.float64), (nf, 3)))
        p = None if pressure is None else np.ascontiguousarray(np.broadcast_to(np.asarray(pressure, dtype=np.float64), (nf,)))
        out = np.zeros(self.n_dofs) if f is None else np.array(f, dtype=np.float64)
        vp = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
        check(lib().jfem_surface_load(self._h, face_type, nf, vp(fc), vp(t), vp(p), int(f is not None), vp(out), 0))
        return out

    def reactions(self, u, f_ext=None):
        uu = np.ascontiguousarray(u, dtype=np.float64)
        fe = None if f_ext is None else np.ascontiguousarray(f_ext, dtype=np.float64)
        la = np.zeros(self.n_dofs)
        check(lib().jfem_reactions(self._h, uu.ctypes.data_as(C.c_void_p), None if fe is None else fe.ctypes.data_as(C.c_void_p),
                                   la.ctypes.data_as(C.c_void_p), 0))
        return la

    def nodal_recover(self, u, field=FIELD_STRESS):
        """(n_nodes, 6) least-squares nodal fit of the Gauss-point strain / stress, order 11 22 33 12 23 13."""
        uu = np.ascontiguousarray(u, dtype=np.float64)
        out = np.zeros((self.n_nodes, 6))
        check(lib().jfem_nodal_recover(self._h, uu.ctypes.data_as(C.c_void_p), field, out.ctypes.data_as(C.c_void_p), 0))
        return out

    def csr_penalty_bc(self, rhs=None, scale=1e10):
        """apply_dirichlet_bc! of the reference's CPU backend on the handle's assembled CSR; returns (penalty, rhs)."""
        r = None if rhs is None else np.array(rhs, dtype=np.float64)
        pen = C.c_double(0.0)
        check(lib().jfem_csr_penalty_bc(self._h, scale, None if r is None else r.ctypes.data_as(C.c_void_p), C.byref(pen), 0))
        return pen.value, r

    def cg(self, b, x0=None, tol=1e-6, relative=False, max_iter=1000, flags=0):
        pb, db, kb = _ptr(b, self.n_dofs)
        if db:
            import torch
            x = torch.zeros_like(b) if x0 is None else x0
        else:
            x = np.zeros(self.n_dofs) if x0 is None else np.array(x0, dtype=np.float64)
        px, dx, kx = _ptr(x, self.n_dofs)
        if dx != db:
            raise TypeError("b and x must both be host arrays or both device tensors")
        it, res = C.c_int(0), C.c_double(0.0)
        check(lib().jfem_cg(self._h, pb, px, tol, int(relative), max_iter, flags, C.byref(it), C.byref(res), db))
        return x, it.value, res.value

    def newton_krylov(self, f_ext, u0=None, newton_tol=1e-6, max_newton=20, max_cg_per_newton=50, forcing_power=0.5,
                      forcing_max=0.9, flags=0):
        pf, df, kf = _ptr(f_ext, self.n_dofs)
        if df:
            import torch
            u = torch.zeros_like(f_ext) if u0 is None else u0
        else:
            u = np.zeros(self.n_dofs) if u0 is None else np.array(u0, dtype=np.float64)
        pu, du, ku = _ptr(u, self.n_dofs)
        ni, ci, res = C.c_int(0), C.c_int(0), C.c_double(0.0)
        cap = max_newton + 1
        hist = np.zeros((cap, 3))
        check(lib().jfem_newton_krylov(self._h, pf, pu, newton_tol, max_newton, max_cg_per_newton, forcing_power, forcing_max,
                                       flags, C.byref(ni), C.byref(ci), C.byref(res), hist.ctypes.data_as(C.c_void_p), cap, df))
        history = [(int(h[0]), float(h[1]), float(h[2])) for h in hist[: ni.value]]
        return u, ni.value, ci.value, res.value, history

    # ---- multi-GPU
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        check(lib().jfem_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, n_ranks, rank, uid: bytes, n_owned_nodes):
        check(lib().jfem_comm_init(self._h, n_ranks, rank, uid, n_owned_nodes))

    def comm_set_halo(self, send: dict, recv: dict):
        nbs = sorted(set(send) | set(recv))
        sp, rp, sn, rn = [0], [0], [], []
        for r in nbs:
            s = np.asarray(send.get(r, []), dtype=np.int32)
            q = np.asarray(recv.get(r, []), dtype=np.int32)
            sn.append(s); rn.append(q)
            sp.append(sp[-1] + s.size); rp.append(rp[-1] + q.size)
        nb = np.asarray(nbs, dtype=np.int32)
        spa, rpa = np.asarray(sp, dtype=np.int64), np.asarray(rp, dtype=np.int64)
        sna = np.ascontiguousarray(np.concatenate(sn) if sn else np.zeros(0), dtype=np.int32)
        rna = np.ascontiguousarray(np.concatenate(rn) if rn else np.zeros(0), dtype=np.int32)
        vp = lambda a: a.ctypes.data_as(C.c_void_p)
        check(lib().jfem_comm_set_halo(self._h, len(nbs), vp(nb), vp(spa), vp(sna), vp(rpa), vp(rna)))
        self._halo_nbs, self._halo_recv_ptr = nbs, rp

    def comm_p2p_export(self) -> bytes:
        buf = C.create_string_buffer(128)
        check(lib().jfem_comm_p2p_export(self._h, buf))
        return buf.raw

    def comm_p2p_seq(self, set_to: int = -1) -> int:
        """Halo sequence number (must agree on all ranks); set_to >= 0 sets it first."""
        out = C.c_int64(0)
        check(lib().jfem_comm_p2p_seq(self._h, set_to, C.byref(out)))
        return int(out.value)

    def comm_p2p_import(self, all_handles: bytes, recv_offsets, halves):
        ro = np.ascontiguousarray(recv_offsets, dtype=np.int64)
        hv = np.ascontiguousarray(halves, dtype=np.int64)
        check(lib().jfem_comm_p2p_import(self._h, all_handles, ro.ctypes.data_as(C.c_void_p), hv.ctypes.data_as(C.c_void_p)))
