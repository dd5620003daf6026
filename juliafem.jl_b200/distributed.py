"""Multi-GPU host plumbing: one process per GPU (torchrun), torch.distributed for the rendezvous, NCCL inside the
library for the data path (halo send/recv + dot-product all-reduce).

Replaces the MPI scaffolding of benchmarks/multigpu_mpi_benchmark.jl:366-446 (every rank builds the mesh, partitions
by contiguous node range, renumbers to local ids, uploads) -- see mesh.partition_mesh for the partition itself.
"""
from __future__ import annotations

import numpy as np

from . import _lib, mesh as _mesh


class PartitionedProblem:
    """The local part of a mesh on this rank + the handle that owns it on the GPU."""

    def __init__(self, m: _mesh.Mesh, rank: int, world: int, device: int, material=(0, (210e9, 0.3)), fixed_dofs=None,
                 options=None, node_offset: int = 0, n_nodes_global: int | None = None, n_elems_global: int | None = None):
        """m: the whole mesh, or (node_offset / n_nodes_global given) the window of it that mesh.lattice_window built for
        this rank.  fixed_dofs: GLOBAL 1-based dof ids (only those of local nodes are kept)."""
        self.rank, self.world = rank, world
        self.elem_type = m.elem_type
        self.n_nodes_global = m.n_nodes if n_nodes_global is None else int(n_nodes_global)
        self.n_elems_global = m.n_elems if n_elems_global is None else int(n_elems_global)
        self.part = _mesh.partition_mesh(m, world, rank, node_offset, n_nodes_global) if world > 1 else None
        if world > 1:
            p = self.part
            self.local_nodes = p.local_nodes
            self.n_owned = p.n_owned
            self.coords_local = np.ascontiguousarray(m.coords[p.local_nodes - 1 - node_offset])
            self.conn_local = p.conn_local
        else:
            self.local_nodes = np.arange(1, m.n_nodes + 1, dtype=np.int64)
            self.n_owned = m.n_nodes
            self.coords_local, self.conn_local = m.coords, m.conn
        self.handle = _lib.Handle(m.elem_type, self.coords_local, self.conn_local, device=device)
        for k, v in (options or {}).items():
            self.handle.set_option(k, v)
        self.handle.set_material(*material)
        self.fixed_local = np.zeros(0, dtype=np.int64)
        if fixed_dofs is not None:
            self.fixed_local = self.to_local_dofs(fixed_dofs)
            self.handle.set_dirichlet(self.fixed_local)

    @classmethod
    def from_lattice(cls, elem_type, dims, box, rank, world, device, fixed_plane_x0=True, **kw):
        """Structured lattice (mesh.hex8_lattice / mesh.tet10_kuhn) partitioned without building the global mesh."""
        w, off, nn, ne = _mesh.lattice_window(elem_type, dims, box, world, rank)
        fixed = None
        if fixed_plane_x0:     # clamp x = 0 (demos/cantilever_physics_gpu.jl:88-93), global dof ids of the window's nodes
            nodes = np.nonzero(np.abs(w.coords[:, 0]) <= 1e-12)[0].astype(np.int64) + 1 + off
            fixed = (3 * (nodes[:, None] - 1) + np.arange(1, 4)[None, :]).ravel()
        return cls(w, rank, world, device, fixed_dofs=fixed, node_offset=off, n_nodes_global=nn, n_elems_global=ne, **kw)

    # ---- numbering helpers (1-based dofs)
    def to_local_dofs(self, gdofs):
        gdofs = np.asarray(gdofs, dtype=np.int64)
        if self.world == 1:
            return gdofs
        node = (gdofs - 1) // 3 + 1
        order = np.argsort(self.local_nodes, kind="stable")
        pos = np.searchsorted(self.local_nodes, node, sorter=order)
        pos = np.minimum(pos, order.size - 1)
        loc = order[pos]
        keep = self.local_nodes[loc] == node
        return 3 * loc[keep] + (gdofs[keep] - 1) % 3 + 1

    def scatter_vector(self, v_global):
        """global dof vector -> local (owned + ghost) vector"""
        if self.world == 1:
            return np.array(v_global, dtype=np.float64)
        return np.asarray(v_global, dtype=np.float64).reshape(-1, 3)[self.local_nodes - 1].ravel().copy()

    def local_global_dofs(self):
        """0-based GLOBAL dof ids of all local dofs (owned first, then ghosts)."""
        return (3 * (self.local_nodes[:, None] - 1) + np.arange(3)[None, :]).ravel()

    def owned_slice(self):
        return slice(0, 3 * self.n_owned)

    def owned_global_dofs(self):
        n = self.local_nodes[: self.n_owned]
        return (3 * (n[:, None] - 1) + np.arange(3)[None, :]).ravel()      # 0-based global dof indices

    # ---- communicator
    def init_comm(self, broadcast_bytes):
        """broadcast_bytes(b: bytes|None) -> bytes : collective that returns rank 0's 128-byte NCCL unique id."""
        if self.world == 1:
            return
        uid = _lib.Handle.comm_unique_id() if self.rank == 0 else None
        uid = broadcast_bytes(uid)
        self.handle.comm_init(self.world, self.rank, uid, self.n_owned)
        self.handle.comm_set_halo(self.part.send, self.part.recv)

    def init_p2p(self, all_gather_object):
        """Switch the halo exchange to direct peer-memory stores (CUDA IPC over NVLink).
        all_gather_object(obj) -> list of every rank's obj (collective)."""
        if self.world == 1:
            return
        h = self.handle
        mine = h.comm_p2p_export()
        row = np.full(self.world, -1, dtype=np.int64)
        for k, r in enumerate(h._halo_nbs):
            row[r] = h._halo_recv_ptr[k]
        half = 3 * h._halo_recv_ptr[-1] + 8
        got = all_gather_object((mine, row, half))
        handles = b"".join(g[0] for g in got)
        offs = np.stack([g[1] for g in got])
        halves = np.array([g[2] for g in got], dtype=np.int64)
        h.comm_p2p_import(handles, offs, halves)


def torch_all_gather_object(dist):
    def gather(obj):
        out = [None] * dist.get_world_size()
        dist.all_gather_object(out, obj)
        return out
    return gather


def torch_broadcast_bytes(dist, device=None):
    """Adapter: broadcast the unique id through an initialised torch.distributed process group."""
    def bcast(b):
        box = [b]
        if device is not None:
            dist.broadcast_object_list(box, src=0, device=device)
        else:
            dist.broadcast_object_list(box, src=0)
        return box[0]
    return bcast
