"""Worker for the 2-rank tests (launched with torch.multiprocessing / torchrun).

mode "gloo": CPU only -- exercises the host-side partition + halo lists + dot all-reduce with the ORACLE doing the
local arithmetic (this is test infrastructure; the product path has no CPU fallback).
mode "nccl": one GPU per rank -- the library's own NCCL halo exchange and distributed CG.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def gloo_worker(rank, world, port, out):
    import torch
    import torch.distributed as dist
    from juliafem.jl_b200 import mesh
    from oracle import oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    m = mesh.tet10_kuhn(4, 3, 3)
    u = mesh.test_vector(m.n_dofs)
    p = mesh.partition_mesh(m, world, rank)
    ul = u.reshape(-1, 3)[p.local_nodes - 1].copy()
    ul[p.n_owned:] = np.nan                                  # ghosts must come from the exchange
    reqs, bufs = [], {}
    for s, ids in p.send.items():
        t = torch.from_numpy(np.ascontiguousarray(ul[ids - 1]))
        reqs.append(dist.isend(t, dst=s))
    for s, ids in p.recv.items():
        bufs[s] = torch.empty((len(ids), 3), dtype=torch.float64)
        reqs.append(dist.irecv(bufs[s], src=s))
    for r in reqs:
        r.wait()
    for s, ids in p.recv.items():
        ul[ids - 1] = bufs[s].numpy()
    assert not np.isnan(ul).any()
    yl = O.matfree(10, m.coords[p.local_nodes - 1], p.conn_local, ul.ravel()).reshape(-1, 3)
    yref = O.matfree(10, m.coords, m.conn, u).reshape(-1, 3)[p.local_nodes[:p.n_owned] - 1]
    err = np.abs(yl[:p.n_owned] - yref).max() / np.abs(yref).max()
    dot = torch.tensor([float((ul[:p.n_owned] * yl[:p.n_owned]).sum())], dtype=torch.float64)
    dist.all_reduce(dot)
    gref = float(u @ O.matfree(10, m.coords, m.conn, u))
    ok = err < 1e-13 and abs(dot.item() - gref) < 1e-12 * abs(gref)
    out[rank] = (ok, err, dot.item(), gref)
    dist.destroy_process_group()


def nccl_main():
    import torch
    import torch.distributed as dist
    from juliafem.jl_b200 import _lib, mesh
    from juliafem.jl_b200.distributed import PartitionedProblem, torch_all_gather_object, torch_broadcast_bytes
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    et = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    m = mesh.tet10_kuhn(16, 6, 12, 4.0, 1.0, 2.0) if et == 10 else mesh.hex8_lattice(9, 8, 13, 0.1)
    fixed = mesh.clamp_dofs(m)
    u = mesh.test_vector(m.n_dofs, fixed)
    b = np.zeros(m.n_dofs); b[2::3] = -1e3
    # single-GPU answers (every rank computes them on its own GPU)
    h1 = _lib.Handle(et, m.coords, m.conn, device=lr)
    h1.set_material(0, (210e9, 0.3)); h1.set_dirichlet(fixed)
    y1 = h1.matvec(u, flags=_lib.PROJECT)
    x1, it1, r1 = h1.cg(b, tol=1e-8, relative=True, max_iter=20000)
    # partitioned
    pp = PartitionedProblem(m, rank, world, lr, fixed_dofs=fixed)
    pp.init_comm(torch_broadcast_bytes(dist, dev))
    if len(sys.argv) > 3 and sys.argv[3].startswith("p2p"):
        pp.init_p2p(torch_all_gather_object(dist))
        if sys.argv[3] == "p2p-unfused":
            pp.handle.set_option("fused_halo", 0)          # separate push / pull kernels instead of the in-kernel exchange
    ul = pp.scatter_vector(u)
    ul[3 * pp.n_owned:] = 1e300                               # ghosts must be overwritten by the halo exchange
    xd = torch.from_numpy(ul).to(dev)
    yd = torch.empty_like(xd)
    pp.handle.matvec(xd, yd, flags=_lib.PROJECT)
    torch.cuda.synchronize()
    yl = yd.cpu().numpy()
    own = pp.owned_global_dofs()
    e_mv = np.abs(yl[pp.owned_slice()] - y1[own]).max() / np.abs(y1).max()
    bl = pp.scatter_vector(b)
    xl, it2, r2 = pp.handle.cg(bl, tol=1e-8, relative=True, max_iter=20000)
    e_cg = np.abs(xl[pp.owned_slice()] - x1[own]).max() / np.abs(x1).max()
    res = torch.tensor([e_mv, e_cg, float(it2), float(it1)], device=dev, dtype=torch.float64)
    dist.all_reduce(res, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("MULTIRANK_RESULT", *[float(v) for v in res.cpu()], flush=True)
    pp.handle.close(); h1.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    if sys.argv[1] == "nccl":
        nccl_main()
