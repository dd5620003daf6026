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
                 options=None):
        self.mesh, self.rank, self.world = m, rank, world
        self.part = _mesh.partition_mesh(m, world, rank) if world > 1 else None
        if world > 1:
            p = self.part
            self.local_nodes = p.local_nodes
            self.n_owned = p.n_owned
            self.handle = _lib.Handle(m.elem_type, m.coords[p.local_nodes - 1], p.conn_local, device=device)
        else:
            self.local_nodes = np.arange(1, m.n_nodes + 1, dtype=np.int64)
            self.n_owned = m.n_nodes
            self.handle = _lib.Handle(m.elem_type, m.coords, m.conn, device=device)
        for k, v in (options or {}).items():
            self.handle.set_option(k, v)
        self.handle.set_material(*material)
        if fixed_dofs is not None:
            self.handle.set_dirichlet(self.to_local_dofs(fixed_dofs))

    # ---- numbering helpers (1-based dofs)
    def to_local_dofs(self, gdofs):
        gdofs = np.asarray(gdofs, dtype=np.int64)
        if self.world == 1:
            return gdofs
        g2l = np.zeros(self.mesh.n_nodes + 1, dtype=np.int64)
        g2l[self.local_nodes] = np.arange(1, self.local_nodes.size + 1)
        node = (gdofs - 1) // 3 + 1
        keep = g2l[node] > 0
        return 3 * (g2l[node[keep]] - 1) + (gdofs[keep] - 1) % 3 + 1

    def scatter_vector(self, v_global):
        """global dof vector -> local (owned + ghost) vector"""
        if self.world == 1:
            return np.array(v_global, dtype=np.float64)
        return np.asarray(v_global, dtype=np.float64).reshape(-1, 3)[self.local_nodes - 1].ravel().copy()

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
