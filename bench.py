#!/usr/bin/env python
"""bench.py -- the driver-facing benchmark of the hot path (BASELINE.json: Tet10 elasticity matvec GDOF/s).

A "step" is ONE matrix-free K.u over the whole mesh (BASELINE.json configs[1]: T1, Tet10 cantilever 88x22x22 cells,
1 075 275 DOF; for N GPUs the box grows to 88x22x(22N) cells and is slab-partitioned by contiguous node ranges, i.e.
weak scaling).  Timed with CUDA events on the stream the kernels are launched on; L2 is flushed between timed steps
(a 512 MiB buffer is overwritten), because the T1 working set (~40 MB) would otherwise sit in the 126 MB L2.
After an untimed clock spin-up (--spinup seconds of the same step, same count on every rank) and W warm-up steps, the K
timed steps (flush, event, K.u, event) are captured once and replayed as ONE CUDA graph, bracketed by barrier +
synchronize, so that no host launch sits between the steps (--graph 0: per-step host launches; automatic fallback if the
capture fails).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload T1|T10|H100|...]

Prints ONE JSON line (rank 0).  `value` = device-resident throughput, `e2e` = same metric through the C ABI with
pinned HOST buffers (H2D + D2H inside the timed region), `roofline` = algorithmic bytes (35 B/DOF Tet10, 35.7 Hex8,
SURVEY.md 8d) / measured step time vs the measured HBM peak, `cpu_baseline` = the CPU oracle timed on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {   # name: (elem_type, cells/nodes, box lengths)
    "T1": (10, (88, 22, 22), (4.0, 1.0, 1.0)),        # 1 075 275 DOF  (configs[1])
    "T10": (10, (192, 48, 48), (4.0, 1.0, 1.0)),      # 10 867 395 DOF
    "TS": (10, (24, 6, 6), (4.0, 1.0, 1.0)),          # small smoke size
    "H100": (8, (321, 321, 321), None),               # 99 228 483 DOF (configs[2]); nodes per direction
    "H12": (8, (161, 161, 161), None),                # 12.5 M DOF
}
BYTES_PER_DOF = {10: 35.0, 8: 35.0 + 2.0 / 3.0, 4: 35.0}
METRIC = "tet10_elasticity_matvec_gdofs"


def build_mesh(name, n_gpus=1):
    from juliafem.jl_b200 import mesh
    et, dims, box = WORKLOADS[name]
    if et == 10:
        cx, cy, cz = dims
        return mesh.tet10_kuhn(cx, cy, cz * n_gpus, box[0], box[1], box[2] * n_gpus)
    nx, ny, nz = dims
    return mesh.hex8_lattice(nx, ny, (nz - 1) * n_gpus + 1, 1.0 / (nx - 1))


def workload_label(name, n_gpus=1):
    """'T1: Tet10 matrix-free K.u, <n> DOF, <m> elements' without building the mesh (same text as the GPU arm's config)."""
    et, dims, _ = WORKLOADS[name]
    if et == 10:
        cx, cy, cz = dims[0], dims[1], dims[2] * n_gpus
        nd, ne = 3 * (2 * cx + 1) * (2 * cy + 1) * (2 * cz + 1), 6 * cx * cy * cz
    else:
        nx, ny, nz = dims[0], dims[1], (dims[2] - 1) * n_gpus + 1
        nd, ne = 3 * nx * ny * nz, (nx - 1) * (ny - 1) * (nz - 1)
    return f"{name}: {'Tet10' if et == 10 else 'Hex8'} matrix-free K.u, {nd} DOF, {ne} elements"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region."""

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index), "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([s.strip() for s in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def cpu_baseline(workload, seconds=12.0):
    """The reference's CPU K.v is a sparse matrix-vector product on the assembled K
    (src/element_assembly_structures.jl:307-309).  Timed with the oracle port on a bounded sample: a sub-box of the
    workload (same element type, spacing, material) small enough that assembly + timing stay within ~10-30 s."""
    from oracle import oracle as O
    from juliafem.jl_b200 import mesh
    et, dims, box = WORKLOADS[workload]
    if et == 10:
        m = mesh.tet10_kuhn(32, 12, 12, 4.0 * 32 / 88, 12 / 22, 12 / 22)
        sample = "Tet10 32x12x12-cell sub-box of the workload (121 875 DOF): assembled CSR SpMV, OpenMP"
    else:
        m = mesh.hex8_lattice(49, 49, 49, 1.0 / 320)
        sample = "Hex8 49^3-node sub-box (352 947 DOF): assembled CSR SpMV, OpenMP"
    t0 = time.perf_counter()
    rp, ci, vals, _ = O.assemble_csr(et, m.coords, m.conn, par=(210e9, 0.3))
    t_asm = time.perf_counter() - t0
    u = mesh.test_vector(m.n_dofs)
    O.spmv(rp, ci, vals, u)
    seconds = float(os.environ.get("JFEM_BENCH_CPU_SECONDS", seconds))     # (the CLI test shortens the timing loop)
    best, t_end, reps = 1e30, time.perf_counter() + min(seconds, 8.0), 0
    while time.perf_counter() < t_end or reps < 3:
        t0 = time.perf_counter()
        O.spmv(rp, ci, vals, u)
        best = min(best, time.perf_counter() - t0)
        reps += 1
    t0 = time.perf_counter()
    O.matfree(et, m.coords, m.conn, u)
    t_mf = time.perf_counter() - t0
    return {"value": m.n_dofs / best / 1e9, "unit": "GDOF/s", "cores": O.num_threads(), "kind": "port", "ms_per_step": best * 1e3,
            "sample": sample + f"; best of {reps}; assembly of the sample took {t_asm:.2f} s ({m.n_elems / t_asm:.0f} elements/s); "
                               f"matrix-free oracle K.u {m.n_dofs / t_mf / 1e9:.4f} GDOF/s"}


def run_reference(args, rank):
    """--impl reference: the reference's own CPU implementation of the path.  The reference is Julia (not installed, and
    the package does not load as shipped), so this arm times the oracle port of its CPU K.v with all host
    threads, rank 0 only."""
    if rank != 0:
        return
    cb = cpu_baseline(args.workload, seconds=10.0)
    out = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "GDOF/s", "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic",
           "config": {"workload": workload_label(args.workload) + " -- CPU arm: assembled CSR K.v of the reference on a bounded sample of this mesh "
                                  "(see cpu_baseline.sample)"},
           "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": "GDOF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="T1", choices=sorted(WORKLOADS))
    ap.add_argument("--patch", type=int, default=0)
    ap.add_argument("--no-flush", action="store_true")
    ap.add_argument("--opt", action="append", default=[], help="library option key=value (jfem_set_option), repeatable")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--spinup", type=float, default=0.8, help="seconds of untimed load before the warm-up steps (clock ramp)")
    ap.add_argument("--graph", type=int, default=-1, help="1: replay the timed loop as one CUDA graph (no host launch jitter between "
                    "ranks); 0: launch every step from the host; default: 1")
    ap.add_argument("--nccl-halo", action="store_true", help="halo through ncclSend/ncclRecv instead of peer-memory stores")
    ap.add_argument("--cg", action="store_true", help="also time a full CG solve (||r|| <= 1e-8 ||b||) on the workload")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    from juliafem.jl_b200 import _lib, mesh

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from juliafem.jl_b200.distributed import PartitionedProblem, torch_all_gather_object, torch_broadcast_bytes
    m = build_mesh(args.workload, world)
    et = m.elem_type
    fixed_g = mesh.clamp_dofs(m)
    total_dofs = m.n_dofs
    pp = PartitionedProblem(m, rank, world, local_rank, material=(_lib.MAT_LINEAR_ELASTIC, (210e9, 0.3)), fixed_dofs=fixed_g,
                            options=dict([("patch_elems", args.patch)] if args.patch else [], **{k: float(v) for k, v in (o.split("=") for o in args.opt)}) or None)
    h = pp.handle
    n_local_dofs, n_own_dofs = 3 * pp.local_nodes.size, 3 * pp.n_owned
    h.set_stream(torch.cuda.current_stream().cuda_stream)
    if world > 1:
        pp.init_comm(torch_broadcast_bytes(dist, dev))
        if not args.nccl_halo:
            pp.init_p2p(torch_all_gather_object(dist))

    u_full = mesh.test_vector(total_dofs, fixed_g)
    u_host = pp.scatter_vector(u_full)
    x = torch.from_numpy(u_host).to(dev)
    y = torch.empty_like(x)
    flush = None if args.no_flush else torch.empty(512 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        h.matvec(x, y, flags=_lib.PROJECT)

    # untimed spin-up: a GPU that has been idle (setup is host work) needs tens of ms under load to reach its boost clock;
    # with W warm-up steps of ~50 us each a rank could still be ramping during the timed region and stall its neighbours
    # (the number of spin-up steps must be the same on every rank: each step is a halo exchange with the neighbours)
    step()                                   # first call builds the patches (host work): keep it out of the step-time estimate
    torch.cuda.synchronize()
    t_spin = time.perf_counter()
    for _ in range(10):
        step()
    torch.cuda.synchronize()
    t_step = torch.tensor([(time.perf_counter() - t_spin) / 10], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_step, op=dist.ReduceOp.MAX)
    n_spin = int(min(20000, max(0, args.spinup / max(float(t_step.item()), 1e-6))))
    for _ in range(n_spin):
        step()
    torch.cuda.synchronize()
    for _ in range(args.warmup):
        step()
    barrier()
    info = h.info()
    launches_before = int(info.total_launches)

    sampler = ClockSampler(local_rank)       # every rank watches its own GPU
    sampler.start()
    use_graph = True if args.graph < 0 else bool(args.graph)

    def timed_loop(ev):
        for k in range(args.steps):
            if flush is not None:
                flush.fill_(float(k))
            ev[k][0].record()
            step()
            ev[k][1].record()

    host_s = None
    if use_graph:
        # The K timed steps (L2 flush, event, K.u, event) are captured once and replayed as ONE graph launch: every rank's
        # GPU then runs its K steps back to back with no host in the loop, so a late host launch on one rank cannot stall
        # its neighbours at the halo gate.  Same kernels, same events, same barrier + synchronize bracket.
        ok = 1.0
        try:
            ev = [(torch.cuda.Event(enable_timing=True, external=True), torch.cuda.Event(enable_timing=True, external=True)) for _ in range(args.steps)]
            g = torch.cuda.CUDAGraph()
            cap = torch.cuda.Stream()
            torch.cuda.synchronize()
            h.set_stream(cap.cuda_stream)          # (synchronises the old stream: must happen outside the capture)
            with torch.cuda.graph(g, stream=cap):
                timed_loop(ev)
        except Exception as exc:                   # capture unsupported here: fall back to per-step host launches
            ok = 0.0
            print(f"[bench] CUDA graph capture failed ({str(exc)[:120]}); timing with per-step launches", file=sys.stderr, flush=True)
        torch.cuda.synchronize()
        h.set_stream(torch.cuda.current_stream().cuda_stream)
        if world > 1:                              # every rank must take the same path
            t = torch.tensor([ok], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            ok = float(t.item())
        if ok:
            barrier()
            g.replay()
            barrier()
        else:
            use_graph = False
            if world > 1:                          # ranks may have enqueued different numbers of exchanges: re-synchronise the halo sequence
                t = torch.tensor([float(h.comm_p2p_seq())], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                h.comm_p2p_seq(int(t.item()) + 2)
    if not use_graph:
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        barrier()
        t0 = time.perf_counter()
        timed_loop(ev)
        host_s = (time.perf_counter() - t0) / args.steps
        barrier()
    launches_timed = int(h.info().total_launches) - launches_before     # kernels of the library launched inside the timed region
    times = np.array([a.elapsed_time(b) for a, b in ev])       # ms, device time of each step
    ms = float(times.mean())
    per_rank = None
    if world > 1:
        st = torch.tensor([ms, float(times.min()), float(np.median(times)), float(times.max())], device=dev, dtype=torch.float64)
        allst = [torch.zeros_like(st) for _ in range(world)]
        dist.all_gather(allst, st)
        per_rank = [[round(float(v), 5) for v in a.cpu()] for a in allst]      # mean, min, median, max of every rank (ms)
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    clocks = sampler.stop()
    if world > 1:
        mine = torch.tensor([clocks["sm_mhz"] or 0.0, float(len(clocks["reasons"]))], device=dev, dtype=torch.float64)
        allc = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allc, mine)
        gathered = [None] * world
        dist.all_gather_object(gathered, clocks["reasons"])
        clocks["per_rank_sm_mhz"] = [float(c[0]) for c in allc]
        vals = [v for v in clocks["per_rank_sm_mhz"] if v > 0]
        clocks["sm_mhz"] = min(vals) if vals else None        # the slowest GPU's median clock under load
        clocks["reasons"] = sorted({r for g in gathered for r in (g or [])})

    # end to end through the C ABI with pinned host buffers (H2D of x, D2H of y inside the timed region)
    xh = torch.empty(n_local_dofs, dtype=torch.float64, pin_memory=True)
    yh = torch.empty(n_local_dofs, dtype=torch.float64, pin_memory=True)
    xh.copy_(torch.from_numpy(u_host))
    xn, yn = xh.numpy(), yh.numpy()
    n_e2e = max(3, min(args.steps, 20))
    for _ in range(2):
        h.matvec(xn, yn, flags=_lib.PROJECT)
    barrier()
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        h.matvec(xn, yn, flags=_lib.PROJECT)
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / n_e2e
    if world > 1:
        t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    checksum = float(np.abs(yn[:n_own_dofs]).sum())

    cg_out = None
    if args.cg:
        # CG time-to-solve (second half of BASELINE.json's metric): uniform body load in -z, clamp x = 0,
        # plain (unpreconditioned) CG exactly as the reference's cg_solve_matfree_gpu!, relative stop 1e-8.
        bfull = np.zeros(total_dofs); bfull[2::3] = -1.0e3
        bd = torch.from_numpy(pp.scatter_vector(bfull)).to(dev)
        xd = torch.zeros_like(bd)
        barrier()
        t0 = time.perf_counter()
        _, cg_it, cg_res = h.cg(bd, x0=xd, tol=1e-8, relative=True, max_iter=200000)
        torch.cuda.synchronize()
        cg_s = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([cg_s], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            cg_s = float(t.item())
        cg_out = {"workload": args.workload, "dofs": total_dofs, "iterations": int(cg_it), "seconds": cg_s, "final_abs_residual": float(cg_res),
                  "tol": "||r|| <= 1e-8 ||b|| (relative; the reference's default is absolute 1e-6)", "ms_per_iteration": 1e3 * cg_s / max(cg_it, 1),
                  "preconditioner": "none (as the reference)"}

    if rank == 0:
        peak, peak_src = peaks()
        gdofs = total_dofs / (ms * 1e-3) / 1e9
        ach = BYTES_PER_DOF[et] * total_dofs / world / (ms * 1e-3) / 1e9     # per GPU
        out = {
            "metric": METRIC, "value": gdofs, "unit": "GDOF/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {'Tet10' if et == 10 else 'Hex8'} matrix-free K.u, {total_dofs} DOF, {m.n_elems} elements, "
                                   f"linear elastic E=210e9 nu=0.3, clamp x=0, deterministic scatter",
                       "l2": "flushed between timed steps (512 MiB write)" if flush is not None else "not flushed",
                       "patch_elems": int(info.patch_elems), "n_patches": int(info.n_patches), "affine_elems": int(info.n_affine_elems), "smem_bytes": int(info.smem_bytes), "blocks_per_sm": int(info.blocks_per_sm), "interface_nodes": int(info.n_interface_nodes),
                       "partition": ("z-slabs by contiguous node range, owner-computes + ghost elements; halo via "
                                     + ("ncclSend/ncclRecv" if args.nccl_halo else "peer-memory stores over NVLink (CUDA IPC) issued from inside the patch kernel; "
                                        "patches that read ghost values run last and wait for the neighbours' flags")) if world > 1 else "single GPU",
                       "ms_min": float(times.min()), "ms_max": float(times.max()), "setup_s": float(info.setup_seconds),
                       "launch": "one CUDA graph replay of the K timed steps" if use_graph else "per-step host launches",
                       "host_enqueue_ms_per_step": None if host_s is None else host_s * 1e3,
                       "per_rank_ms_mean_min_median_max": per_rank, "rank0_first_steps_ms": [round(float(v), 4) for v in times[:16]]},
            "e2e": {"value": total_dofs / e2e_s / 1e9, "unit": "GDOF/s", "h2d_bytes_per_step": 8 * n_local_dofs, "d2h_bytes_per_step": 8 * n_local_dofs,
                    "ms_per_step": e2e_s * 1e3, "checksum_abs_y": checksum},
            "gpu_launches": launches_timed,
            "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                         "peak_source": peak_src, "bytes_per_dof": BYTES_PER_DOF[et],
                         "note": "achieved = algorithmic bytes of one K.u / CUDA-event time of the whole step (patch kernel + interface reduce)"},
            "clocks": clocks,
        }
        prof = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(prof):
            try:
                out["roofline"]["traffic"] = json.load(open(prof)).get(args.workload)
            except Exception:
                pass
        if cg_out is not None:
            out["cg_time_to_solve"] = cg_out
        if world == 1 and not args.no_cpu:
            out["cpu_baseline"] = cpu_baseline(args.workload)
        print(json.dumps(out), flush=True)
    h.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
