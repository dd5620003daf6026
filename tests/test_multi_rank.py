"""N > 1 path: world_size-2 gloo test on CPU (host logic) and a 2-GPU NCCL test (library data path)."""
import os
import socket
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.parametrize("world", [2, 3])          # 3: the middle rank exchanges with two neighbours
def test_gloo_world2_halo_and_allreduce(oracle, world):
    import torch.multiprocessing as mp
    sys.path.insert(0, HERE)
    import multirank_worker as W
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        out = mgr.dict()
        port = _free_port()
        procs = [ctx.Process(target=W.gloo_worker, args=(r, world, port, out)) for r in range(world)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(120)
            assert p.exitcode == 0
        assert len(out) == world and all(v[0] for v in out.values()), dict(out)


@pytest.mark.gpu
@pytest.mark.parametrize("halo", ["nccl", "p2p", "p2p-unfused"])
@pytest.mark.parametrize("et", [10, 8])
def test_nccl_two_gpus_matvec_and_cg(et, halo):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(HERE, "multirank_worker.py"), "nccl", str(et), halo]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    line = [l for l in r.stdout.splitlines() if l.startswith("MULTIRANK_RESULT")]
    assert r.returncode == 0 and line, r.stdout[-2000:] + r.stderr[-2000:]
    e_mv, e_cg, it2, it1 = [float(v) for v in line[0].split()[1:]]
    assert e_mv < 1e-12          # owned rows of K.u identical to the single-GPU result (p2p + Tet10: halo fused into the patch kernel)
    assert e_cg < 1e-6 and abs(it2 - it1) <= max(3, it1 // 50)
