#!/usr/bin/env python
"""bench.py -- the driver-facing benchmark of the hot path (BASELINE.json: Tet10 elasticity matvec GDOF/s at 1/2/4/8 B200;
CG time-to-solve at 10 M DOF).

A "step" is ONE matrix-free K.u over the whole mesh (BASELINE.json configs[1]: T1, Tet10 cantilever 88x22x22 cells,
1 075 275 DOF; for N GPUs the box grows to 88x22x(22N) cells and is slab-partitioned by contiguous node ranges, i.e.
weak scaling).  Timed with CUDA events on the stream the kernels are launched on; L2 is flushed between timed steps
(a 512 MiB buffer is overwritten), because the T1 working set (~40 MB) would otherwise sit in the 126 MB L2.
After an untimed clock spin-up (--spinup seconds of the same step, same count on every rank) and W warm-up steps, the K
timed steps (flush, event, K.u, event) are captured once and replayed as ONE CUDA graph, bracketed by barrier +
synchronize, so that no host launch sits between the steps (--graph 0: per-step host launches; automatic fallback if the
capture fails).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload T1|T10|H100|...]

Prints ONE JSON line (rank 0).  Besides the contract keys:
  parity            BEFORE timing, the GPU K.u of the very mesh that is timed is compared with the CPU oracle's matrix-free
                    K.u (whole mesh when it is small enough, otherwise all rows of a node sample that always contains the
                    partition-interface layer); every rank checks its own rows; the run FAILS above 1e-12 (north_star).
  e2e               same metric through the C ABI with pinned HOST buffers (H2D + D2H inside the timed region)
  roofline          algorithmic bytes (35 B/DOF Tet10, 35.7 Hex8) / measured step time vs the measured HBM peak
  cpu_baseline      the reference's CPU K.v (assembled CSR SpMV, oracle port) on the SAME mesh when it fits, all host threads
  cg_time_to_solve  plain CG to ||r|| <= 1e-8 ||b|| on the 10.9 M-DOF Tet10 cantilever (N = 1) / on the workload itself (N > 1);
                    cg_time_to_solve_block_jacobi: the same solve with the opt-in 3x3 block-Jacobi preconditioner (N = 1)
  assembly          coloured CSR assembly throughput at 2.5 M Tet10 elements (elements/s, fraction of the 3.7 KB/element HBM
                    roofline), pattern build time, CSR SpMV
  plasticity        BASELINE.json configs[4]: J2 plasticity, ~30 % of the Gauss points yielding: state update + assembled tangent
  neo_hookean       BASELINE.json configs[3]: Neo-Hookean residual / tangent K(u).v at 10.9 M DOF + a bounded Newton-Krylov sample
  config.clock_ramp before timing every rank is kept busy until NVML reports >= 90 % of the maximum SM clock on ALL ranks (an idle
                    B200 sits at 120 MHz); what was seen is reported here
  linear_static     BASELINE.json configs[0]: examples/linear_static.jl end to end on the GPU against the value the reference holds
                    (max |u| = 2.4052929896922337) + the CPU assemble / direct-solve time of the same problem
  hex8_weak         BASELINE.json configs[2]: Hex8 lattice, 12.5 M DOF per GPU (99.6 M DOF at N = 8), matrix-free K.u, with its
                    own parity check -- the per-N values give the weak-scaling efficiency of the 100 M-DOF target
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
    "HS": (8, (25, 25, 25), None),                    # small smoke size
}
BYTES_PER_DOF = {10: 35.0, 8: 35.0 + 2.0 / 3.0, 4: 35.0}
ASSEMBLY_BYTES_PER_ELEM = 3.7e3       # values written (2.8 KB) + state + connectivity/coordinates, Tet10
FP64_FLOP_PER_ELEM = {10: 343 * 2 + 170.0, 8: 1400.0}   # SASS of the closed forms: DFMA counted twice (Tet10 affine: 343 DFMA + 170 DADD/DMUL)
METRIC = "tet10_elasticity_matvec_gdofs"
PARITY_TOL = 1e-12
MAT = (210e9, 0.3)


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def lattice_args(name, n_gpus=1):
    """(elem_type, dims, box) of the workload weak-scaled to n_gpus (the box grows in z)."""
    et, dims, box = WORKLOADS[name]
    if et == 10:
        cx, cy, cz = dims
        return et, (cx, cy, cz * n_gpus), (box[0], box[1], box[2] * n_gpus)
    nx, ny, nz = dims
    return et, (nx, ny, (nz - 1) * n_gpus + 1), 1.0 / (nx - 1)


def build_mesh(name, n_gpus=1):
    from juliafem.jl_b200 import mesh
    et, dims, box = lattice_args(name, n_gpus)
    if et == 10:
        return mesh.tet10_kuhn(dims[0], dims[1], dims[2], box[0], box[1], box[2])
    return mesh.hex8_lattice(dims[0], dims[1], dims[2], box)


def workload_label(name, n_gpus=1):
    """'T1: Tet10 matrix-free K.u, <n> DOF, <m> elements' without building the mesh (same text as the GPU arm's config)."""
    et, dims, _ = lattice_args(name, n_gpus)
    if et == 10:
        cx, cy, cz = dims
        nd, ne = 3 * (2 * cx + 1) * (2 * cy + 1) * (2 * cz + 1), 6 * cx * cy * cz
    else:
        nx, ny, nz = dims
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
    """nvidia-smi clocks/throttle reasons sampled DURING the spin-up + timed region."""

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index), "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [s.strip() for s in line.split(",")]))

    def stop(self, t_from=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t_from is None or t >= t_from] or [r for _, r in self.rows[-3:]]
        sm = [float(r[0]) for r in rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


class NvmlClock:
    """Instantaneous SM clock of one GPU through NVML (nvidia_ml_py), addressed by the UUID / PCI bus id torch reports so that
    CUDA_VISIBLE_DEVICES renumbering cannot point it at another GPU."""

    def __init__(self, props, index):
        import pynvml
        self.nv = pynvml
        pynvml.nvmlInit()
        self.h = None
        for make in (lambda: pynvml.nvmlDeviceGetHandleByUUID("GPU-" + str(props.uuid)),
                     lambda: pynvml.nvmlDeviceGetHandleByPciBusId(f"{props.pci_domain_id:08X}:{props.pci_bus_id:02X}:{props.pci_device_id:02X}.0"),
                     lambda: pynvml.nvmlDeviceGetHandleByIndex(index)):
            try:
                self.h = make()
                break
            except Exception:
                continue
        if self.h is None:
            raise RuntimeError("no NVML handle")

    def read(self):
        """(current SM MHz, maximum SM MHz)"""
        return (float(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)),
                float(self.nv.nvmlDeviceGetMaxClockInfo(self.h, self.nv.NVML_CLOCK_SM)))


def wait_for_clocks(step, sync, read_clock, allmin, allmax, chunk=200, max_seconds=3.0, frac=0.9):
    """Keep every GPU busy (chunks of `chunk` steps, the SAME count on every rank: a step is a halo exchange) until the SM
    clock of EVERY rank has reached frac x its maximum, or max_seconds have passed.  An idle B200 sits at 120 MHz; a rank that
    has not ramped up when the timed region starts shows up as a slow rank that every neighbour then waits for.
    The clock is read while the chunk is still executing (the steps are only enqueued), i.e. under load.
    read_clock() -> (mhz, max_mhz) or raises; allmin / allmax reduce a float over the ranks (collective), which also keeps
    the loop count identical everywhere.  Returns a small report for the JSON line."""
    t0 = time.perf_counter()
    rounds, mhz, mx, all_ok = 0, None, None, 1.0
    while True:
        for _ in range(chunk):
            step()
        try:
            mhz, mx = read_clock()
            ok = 1.0 if (mx <= 0 or mhz >= frac * mx) else 0.0
        except Exception:
            mhz, mx, ok = None, None, 1.0              # no NVML: nothing to wait for
        sync()
        rounds += 1
        all_ok = allmin(ok)
        elapsed = allmax(time.perf_counter() - t0)
        if all_ok >= 1.0 or elapsed > max_seconds or rounds >= 1000:
            break
    return {"rounds": rounds, "steps": rounds * chunk, "sm_mhz_under_load": mhz, "sm_max_mhz": mx,
            "all_ranks_ramped": bool(all_ok >= 1.0), "seconds": time.perf_counter() - t0}


# ------------------------------------------------------------------------------------------------ CPU side (oracle)

def cpu_baseline(workload, seconds=12.0, threads=None):
    """The reference's CPU K.v is a sparse matrix-vector product on the assembled K
    (src/element_assembly_structures.jl:307-309).  Timed with the oracle port on all host threads: on the workload's own
    mesh when its CSR fits the time budget (<= 1.6 M DOF: T1 is ~92 M non-zeros), otherwise on a sub-box of it."""
    from oracle import oracle as O
    from juliafem.jl_b200 import mesh
    nthr = O.set_num_threads(threads or host_threads())
    et, dims, box = WORKLOADS[workload]
    nd = int(workload_label(workload).split(", ")[1].split()[0])
    same = nd <= 1_600_000
    if same:
        m = build_mesh(workload)
        sample = f"the workload's own mesh ({m.n_dofs} DOF): assembled CSR SpMV, OpenMP"
    elif et == 10:
        m = mesh.tet10_kuhn(32, 12, 12, 4.0 * 32 / 88, 12 / 22, 12 / 22)
        sample = "Tet10 32x12x12-cell sub-box of the workload (121 875 DOF): assembled CSR SpMV, OpenMP"
    else:
        m = mesh.hex8_lattice(49, 49, 49, 1.0 / 320)
        sample = "Hex8 49^3-node sub-box (352 947 DOF): assembled CSR SpMV, OpenMP"
    if os.environ.get("JFEM_BENCH_CPU_SMALL"):           # (the CLI test keeps the reference arm short)
        m, same = mesh.tet10_kuhn(12, 4, 4, 1.5, 0.5, 0.5), False
        sample = "Tet10 12x4x4-cell sub-box (test mode): assembled CSR SpMV, OpenMP"
    t0 = time.perf_counter()
    rp, ci, vals, _ = O.assemble_csr(m.elem_type, m.coords, m.conn, par=MAT)
    t_asm = time.perf_counter() - t0
    u = mesh.test_vector(m.n_dofs)
    O.spmv(rp, ci, vals, u)
    seconds = float(os.environ.get("JFEM_BENCH_CPU_SECONDS", seconds))
    best, t_end, reps = 1e30, time.perf_counter() + min(seconds, 8.0), 0
    while time.perf_counter() < t_end or reps < 3:
        t0 = time.perf_counter()
        O.spmv(rp, ci, vals, u)
        best = min(best, time.perf_counter() - t0)
        reps += 1
    return {"value": m.n_dofs / best / 1e9, "unit": "GDOF/s", "cores": nthr, "kind": "port", "ms_per_step": best * 1e3, "same_mesh": bool(same),
            "sample": sample + f"; best of {reps}; {nthr} OpenMP threads; assembly of that mesh took {t_asm:.2f} s "
                               f"({m.n_elems / t_asm:.0f} elements/s)"}


def run_reference(args, rank):
    """--impl reference: the reference's own CPU implementation of the path.  The reference is Julia (not installed, and
    the package does not load as shipped), so this arm times the oracle port of its CPU K.v with all host
    threads, rank 0 only (under torchrun the other ranks exit; OMP_NUM_THREADS=1 set by torchrun is overridden)."""
    if rank != 0:
        return
    os.environ["OMP_NUM_THREADS"] = str(host_threads())
    cb = cpu_baseline(args.workload, seconds=10.0)
    out = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "GDOF/s", "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic",
           "config": {"workload": workload_label(args.workload, args.gpus) + " -- CPU arm: assembled CSR K.v of the reference (oracle port), "
                                  + ("on this very mesh" if (cb["same_mesh"] and args.gpus == 1) else
                                     "on a bounded sample of this mesh (see cpu_baseline.sample; GDOF/s is DOF-normalised)"),
                      "same_mesh": bool(cb["same_mesh"] and args.gpus == 1), "omp_threads": cb["cores"]},
           "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": "GDOF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


def parity_check(pp, y_local, u_local, world, full_limit=1_600_000, sample_nodes=12000, seed=7):
    """GPU K.u (projected) of this rank's mesh against the oracle's matrix-free K.u on the same local mesh (owned +
    ghost elements: the owned rows are complete).  Small meshes: every owned row.  Large meshes: all rows of a node
    sample = every owned node of an element that touches a ghost node (the halo path) capped at half the sample, plus
    random owned nodes; only the elements touching the sample are integrated.  Returns (max rel err, #dofs, seconds)."""
    from oracle import oracle as O
    O.set_num_threads(max(1, host_threads() // world))
    t0 = time.perf_counter()
    n_owned = pp.n_owned
    conn = pp.conn_local
    et = pp.elem_type
    if 3 * n_owned <= full_limit:
        nodes = np.arange(1, n_owned + 1)
        sub = conn
    else:
        rng = np.random.default_rng(seed + pp.rank)
        ghosty = (conn > n_owned).any(axis=1)
        near = np.unique(conn[ghosty])
        near = near[near <= n_owned]
        if near.size > sample_nodes // 2:
            near = rng.choice(near, sample_nodes // 2, replace=False)
        rest = rng.choice(n_owned, min(n_owned, sample_nodes - near.size), replace=False) + 1
        nodes = np.unique(np.concatenate([near, rest]))
        mark = np.zeros(pp.local_nodes.size + 1, dtype=bool)
        mark[nodes] = True
        sub = conn[mark[conn].any(axis=1)]
    ref = O.matfree(et, pp.coords_local, sub, u_local, par=MAT, fixed_dofs=pp.fixed_local if pp.fixed_local.size else None)
    rows = (3 * (nodes[:, None] - 1) + np.arange(3)[None, :]).ravel()
    scale = np.abs(ref[: 3 * n_owned]).max()
    err = float(np.abs(y_local[rows] - ref[rows]).max() / (scale if scale > 0 else 1.0))
    return err, int(rows.size), time.perf_counter() - t0


# ------------------------------------------------------------------------------------------------ GPU side

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
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU legs (cpu_baseline AND the oracle parity check)")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--spinup", type=float, default=0.8, help="seconds of untimed load before the warm-up steps (clock ramp)")
    ap.add_argument("--graph", type=int, default=-1, help="1: replay the timed loop as one CUDA graph (no host launch jitter between "
                    "ranks); 0: launch every step from the host; default: 1")
    ap.add_argument("--nccl-halo", action="store_true", help="halo through ncclSend/ncclRecv instead of peer-memory stores")
    ap.add_argument("--cg", default="auto", help="CG time-to-solve: workload name, 'same', 'none' or 'auto' (T10 at N=1, 'same' at N>1)")
    ap.add_argument("--hex8", default="auto", help="secondary Hex8 weak-scaling measurement: workload name per GPU, 'none' or 'auto' (H12)")
    ap.add_argument("--assembly", default="P10", choices=["P10", "small", "none"], help="assembled-path legs (N = 1): mesh size")
    ap.add_argument("--neohooke", type=int, default=1, help="1: Neo-Hookean sample on the CG mesh (N = 1)")
    ap.add_argument("--linear-static", type=int, default=1, help="1: the reference's examples/linear_static.jl end to end against its own expected value (N = 1)")
    ap.add_argument("--jacobi", type=int, default=1, help="1: repeat the CG time-to-solve with the opt-in block-Jacobi preconditioner (N = 1)")
    ap.add_argument("--no-extras", action="store_true", help="skip cg / neo_hookean / assembly / plasticity / hex8_weak (kernel timing only)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    os.environ["OMP_NUM_THREADS"] = str(max(1, host_threads() // world))   # torchrun sets 1: the patch builder and the oracle use OpenMP

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
    sampler = ClockSampler(local_rank)       # every rank watches its own GPU, from before the spin-up on
    sampler.start()

    from juliafem.jl_b200.distributed import PartitionedProblem, torch_all_gather_object, torch_broadcast_bytes
    lib_opts = dict([("patch_elems", args.patch)] if args.patch else [], **{k: float(v) for k, v in (o.split("=") for o in args.opt)}) or None
    do_cpu = not args.no_cpu
    flush = None if args.no_flush else torch.empty(512 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(v):
        if world == 1:
            return float(v)
        t = torch.tensor([float(v)], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allmin(v):
        if world == 1:
            return float(v)
        t = torch.tensor([float(v)], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return float(t.item())

    try:
        nvml_clock = NvmlClock(torch.cuda.get_device_properties(local_rank), local_rank)
    except Exception:
        nvml_clock = None

    def read_clock():
        if nvml_clock is None:
            raise RuntimeError("NVML unavailable")
        return nvml_clock.read()

    def setup_problem(workload):
        et, dims, box = lattice_args(workload, world)
        pp = PartitionedProblem.from_lattice(et, dims, box, rank, world, local_rank, material=(_lib.MAT_LINEAR_ELASTIC, MAT), options=lib_opts)
        pp.handle.set_stream(torch.cuda.current_stream().cuda_stream)
        if world > 1:
            pp.init_comm(torch_broadcast_bytes(dist, dev))
            if not args.nccl_halo:
                pp.init_p2p(torch_all_gather_object(dist))
        gd = pp.local_global_dofs()
        u_host = mesh.hashed_vector(gd)
        if pp.fixed_local.size:
            u_host[pp.fixed_local - 1] = 0.0
        return pp, u_host

    def check_parity(pp, u_host, x, y):
        """K.u of this mesh against the oracle, all ranks; returns the dict for the JSON line (rank 0 aggregates)."""
        pp.handle.matvec(x, y, flags=_lib.PROJECT)
        torch.cuda.synchronize()
        err, ndofs, secs = parity_check(pp, y.cpu().numpy(), u_host, world)
        worst = allmax(err)
        secs = allmax(secs)
        tot = ndofs
        if world > 1:
            t = torch.tensor([float(ndofs)], device=dev, dtype=torch.float64)
            dist.all_reduce(t)
            tot = int(t.item())
        return {"rel_err": worst, "tol": PARITY_TOL, "ok": bool(worst <= PARITY_TOL), "checked_dofs": tot, "ranks_checked": world,
                "oracle": "matrix-free K.u of the CPU oracle (oracle/jfem_oracle.c) on every rank's local mesh, same vector, Dirichlet rows zeroed",
                "oracle_seconds": secs}

    def time_matvec(pp, x, y, steps, warmup, spinup):
        """W warm-up + K timed K.u (L2 flushed in between), as one CUDA graph when possible.  Returns per-step ms array, info."""
        h = pp.handle

        def step():
            h.matvec(x, y, flags=_lib.PROJECT)

        step()
        torch.cuda.synchronize()
        t_spin = time.perf_counter()
        for _ in range(10):
            step()
        torch.cuda.synchronize()
        t_step = allmax((time.perf_counter() - t_spin) / 10)
        n_spin = int(min(20000, max(0, spinup / max(t_step, 1e-6))))     # same count on every rank: each step is a halo exchange
        for _ in range(n_spin):
            step()
        torch.cuda.synchronize()
        ramp = wait_for_clocks(step, torch.cuda.synchronize, read_clock, allmin, allmax, chunk=max(50, min(2000, n_spin // 4)),
                               max_seconds=3.0) if spinup > 0 else None
        for _ in range(warmup):
            step()
        barrier()
        launches_before = int(h.info().total_launches)
        t_timed = time.perf_counter()
        use_graph = True if args.graph < 0 else bool(args.graph)

        def timed_loop(ev):
            for k in range(steps):
                if flush is not None:
                    flush.fill_(float(k))
                ev[k][0].record()
                step()
                ev[k][1].record()

        host_s = None
        if use_graph:
            ok = 1.0
            try:
                ev = [(torch.cuda.Event(enable_timing=True, external=True), torch.cuda.Event(enable_timing=True, external=True)) for _ in range(steps)]
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
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
            barrier()
            t0 = time.perf_counter()
            timed_loop(ev)
            host_s = (time.perf_counter() - t0) / steps
            barrier()
        launches = int(h.info().total_launches) - launches_before
        times = np.array([a.elapsed_time(b) for a, b in ev])
        return times, {"use_graph": use_graph, "host_s": host_s, "launches": launches, "t_timed": t_timed, "clock_ramp": ramp}

    # ================================================================ main workload
    pp, u_host = setup_problem(args.workload)
    h = pp.handle
    et = pp.elem_type
    total_dofs, total_elems = 3 * pp.n_nodes_global, pp.n_elems_global
    n_local_dofs, n_own_dofs = 3 * pp.local_nodes.size, 3 * pp.n_owned
    x = torch.from_numpy(u_host).to(dev)
    y = torch.empty_like(x)
    parity = None
    if do_cpu and not args.no_parity:
        parity = check_parity(pp, u_host, x, y)
        if not parity["ok"] and world > 1 and not args.nccl_halo:
            # The halo fused into the patch kernel failed the oracle check on this partition: say so in the line and measure
            # with the separate push / pull kernels instead (the decision is collective: parity["ok"] is a max over the ranks).
            first = {"rel_err": parity["rel_err"], "path": "halo fused into the patch kernel"}
            h.set_option("fused_halo", 0)
            parity = check_parity(pp, u_host, x, y)
            parity["first_attempt_failed"] = first
            parity["path"] = "fused_halo = 0 (separate push / pull kernels) after the fused path failed the check"
        if not parity["ok"]:
            if rank == 0:
                print(json.dumps({"metric": METRIC, "error": "parity check failed", "parity": parity}), flush=True)
            raise SystemExit(3)
    times, tinfo = time_matvec(pp, x, y, args.steps, args.warmup, args.spinup)
    info = h.info()
    ms = float(times.mean())
    per_rank = None
    if world > 1:
        st = torch.tensor([ms, float(times.min()), float(np.median(times)), float(times.max())], device=dev, dtype=torch.float64)
        allst = [torch.zeros_like(st) for _ in range(world)]
        dist.all_gather(allst, st)
        per_rank = [[round(float(v), 5) for v in a.cpu()] for a in allst]      # mean, min, median, max of every rank (ms)
        ms = allmax(ms)
    clocks = sampler.stop(tinfo["t_timed"])
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
    e2e_s = allmax((time.perf_counter() - t0) / n_e2e)
    checksum = float(np.abs(yn[:n_own_dofs]).sum())
    if world > 1:
        t = torch.tensor([checksum], device=dev, dtype=torch.float64)
        dist.all_reduce(t)
        checksum = float(t.item())

    def cg_solve(pp_, label, flags=0):
        """plain (unpreconditioned) CG exactly as the reference's cg_solve_matfree_gpu!, relative stop 1e-8; uniform body
        load in -z lumped to the nodes, clamp x = 0.  flags = JACOBI: the opt-in 3x3 block-Jacobi PCG, same stop."""
        hh = pp_.handle
        b = np.zeros(3 * pp_.local_nodes.size)
        b[2::3] = -1.0e3
        bd = torch.from_numpy(b).to(dev)
        xd = torch.zeros_like(bd)
        hh.matvec(xd, torch.empty_like(xd), flags=_lib.PROJECT)    # (patch build outside the timing)
        barrier()
        t0 = time.perf_counter()
        _, cg_it, cg_res = hh.cg(bd, x0=xd, tol=1e-8, relative=True, max_iter=200000, flags=flags)
        torch.cuda.synchronize()
        cg_s = allmax(time.perf_counter() - t0)
        return {"workload": label, "dofs": 3 * pp_.n_nodes_global, "iterations": int(cg_it), "seconds": cg_s, "final_abs_residual": float(cg_res),
                "converged": bool(cg_it < 200000), "tol": "||r|| <= 1e-8 ||b|| (relative; the reference's default is absolute 1e-6)",
                "ms_per_iteration": 1e3 * cg_s / max(cg_it, 1),
                "preconditioner": "3x3 block-Jacobi (opt-in, not in the reference; the time includes building the blocks)" if flags & _lib.JACOBI
                                  else "none (as the reference)",
                "setup_s": float(hh.info().setup_seconds), "gdof_iterations_per_s": 3 * pp_.n_nodes_global * cg_it / cg_s / 1e9}

    def guarded(fn, pair=False):
        """the secondary legs must never take the headline line down with them: a failure is reported in place"""
        try:
            return fn()
        except Exception as exc:   # noqa: BLE001
            err = {"error": f"{type(exc).__name__}: {str(exc)[:300]}"}
            print(f"[bench] secondary leg failed: {err['error']}", file=sys.stderr, flush=True)
            return (err, None) if pair else err

    def timed_calls(fn, reps=5, warm=1):
        """CUDA-event time (ms, mean) of fn(), L2 flushed before every repetition"""
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        ts = []
        for k in range(reps):
            if flush is not None:
                flush.fill_(float(k))
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return float(np.mean(ts))

    def neo_hookean_leg(pp_, label):
        """BASELINE.json configs[3]: Neo-Hookean finite-strain cantilever at ~10 M DOF, on the mesh of the CG leg: residual
        f_int(u), matrix-free tangent K(u).v, and a BOUNDED Newton-Krylov sample (one Newton step, CG capped)."""
        hh = pp_.handle
        hh.set_material(_lib.MAT_NEO_HOOKEAN, (3.0e6, 0.45))
        nd = 3 * pp_.local_nodes.size
        G = 0.05 * np.random.default_rng(1).standard_normal((3, 3))
        u = torch.from_numpy(np.ascontiguousarray((pp_.coords_local @ G.T).ravel())).to(dev)
        f = torch.empty_like(u)
        v = torch.from_numpy(mesh.test_vector(nd, pp_.fixed_local if pp_.fixed_local.size else None)).to(dev)
        hh.internal_force(u, f)
        torch.cuda.synchronize()
        ms_r = timed_calls(lambda: hh.internal_force(u, f))
        hh.set_linearization(u)
        ms_t = timed_calls(lambda: hh.matvec(v, f, flags=_lib.TANGENT | _lib.PROJECT))
        top = np.nonzero(np.abs(pp_.coords_local[:, 2] - pp_.coords_local[:, 2].max()) < 1e-12)[0]
        fext = np.zeros(nd)
        fext[3 * top + 2] = -3.0e3 / top.size
        fd = torch.from_numpy(fext).to(dev)
        cap = 300
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        _, nit, cgit, res, hist = hh.newton_krylov(fd, newton_tol=1e-6, max_newton=1, max_cg_per_newton=cap, forcing_max=1e-3)
        torch.cuda.synchronize()
        t_nk = time.perf_counter() - t0
        return {"workload": label + " -- Neo-Hookean (E = 3e6, nu = 0.45), BASELINE.json configs[3]", "dofs": nd,
                "residual_ms": ms_r, "residual_gdofs": nd / ms_r / 1e6, "tangent_matvec_ms": ms_t, "tangent_gdofs": nd / ms_t / 1e6,
                "newton_krylov_sample": {"newton_steps": int(nit), "cg_iterations": int(cgit), "cg_cap_per_newton": cap, "seconds": t_nk,
                                         "ms_per_cg_iteration": 1e3 * (t_nk - 2e-3 * ms_r) / max(int(cgit), 1), "residual_after": float(res),
                                         "note": "bounded sample: ONE Newton step with the tangent CG capped (a converged solve at this size "
                                                 "takes tens of thousands of CG iterations); the per-iteration cost is what scales"},
                "l2": "flushed between repetitions" if flush is not None else "not flushed", "setup_s": float(hh.info().setup_seconds)}

    def assembly_legs(size):
        """Assembled path at P10 scale (Tet10 block 75^3 cells, 2 531 250 elements, 10.3 M DOF; 'small': 64x16x16 cells):
        (a) linear-elastic coloured CSR assembly = `assembly`; (b) BASELINE.json configs[4]: J2 plasticity with ~30 % of the
        Gauss points yielding -- integration-point state update (internal force) + assembled consistent tangent."""
        peak, _ = peaks()
        ma = mesh.tet10_kuhn(75, 75, 75, 1.0, 1.0, 1.0) if size == "P10" else mesh.tet10_kuhn(64, 16, 16, 4.0, 1.0, 1.0)
        ha = _lib.Handle(10, ma.coords, ma.conn, device=local_rank)
        try:
            for k_, v_ in (lib_opts or {}).items():
                ha.set_option(k_, v_)
            ha.set_material(_lib.MAT_LINEAR_ELASTIC, MAT)
            ha.set_stream(torch.cuda.current_stream().cuda_stream)
            ua = torch.zeros(ma.n_dofs, dtype=torch.float64, device=dev)
            ha.matvec(ua, torch.empty_like(ua))            # patch build (needed for the internal element order) outside the pattern timing
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            _, nnz = ha.csr_size()                         # node adjacency + colouring on the host, pattern expansion on the device
            t_pat = time.perf_counter() - t0
            ms_asm = timed_calls(lambda: ha.assemble_csr(ua), reps=4)
            v = torch.from_numpy(mesh.test_vector(ma.n_dofs)).to(dev)
            y1, y2 = torch.empty_like(v), torch.empty_like(v)
            ms_spmv = timed_calls(lambda: ha.spmv(v, y1), reps=4)
            ha.matvec(v, y2)
            torch.cuda.synchronize()
            asm = {"workload": f"Tet10 block, {ma.n_elems} elements, {ma.n_dofs} DOF, linear elastic, coloured scatter into the reference's CSR pattern",
                   "elements_per_s": ma.n_elems / (ms_asm * 1e-3), "ms": ms_asm, "pattern_build_s": t_pat, "nnz": nnz,
                   "bytes_per_element": ASSEMBLY_BYTES_PER_ELEM, "frac_hbm_roofline": ASSEMBLY_BYTES_PER_ELEM * ma.n_elems / (ms_asm * 1e-3) / 1e9 / peak,
                   "spmv_ms": ms_spmv, "spmv_gdofs": ma.n_dofs / ms_spmv / 1e6, "spmv_GBs": 12.0 * nnz / ms_spmv / 1e6,
                   "spmv_frac_hbm": 12.0 * nnz / ms_spmv / 1e6 / peak, "spmv_vs_matfree_rel": float((y1 - y2).abs().max() / y2.abs().max()),
                   "kernel": "one thread per (element, column) (elem_columns_kernel)" if (lib_opts or {}).get("assembly_kernel", 1) == 0
                             else "one warp per element, geometry shared by the 30 columns (elem_warp_kernel)",
                   "l2": "flushed between repetitions" if flush is not None else "not flushed"}
            del y1, y2
            # ---- configs[4]: uniaxial strain growing along x, so that the points with x > 0.7 yield (2 mu eps_xx > sigma_y)
            par = (200e9, 0.3, 100e6, 10e9)
            ha.set_material(_lib.MAT_PERFECT_PLASTICITY, par)
            e_y = par[2] / (2.0 * par[0] / (2.0 * (1.0 + par[1])))
            uh = np.zeros((ma.n_nodes, 3))
            uh[:, 0] = e_y * ma.coords[:, 0] ** 2 / 1.4
            up = torch.from_numpy(uh.ravel()).to(dev)
            fp = torch.empty_like(up)
            ha.internal_force(up, fp)
            torch.cuda.synchronize()
            ms_state = timed_calls(lambda: ha.internal_force(up, fp), reps=4)
            ms_tan = timed_calls(lambda: ha.assemble_csr(up), reps=4)
            ha.set_linearization(up)
            y1, y2 = torch.empty_like(v), torch.empty_like(v)
            ha.spmv(v, y1); ha.matvec(v, y2, flags=_lib.TANGENT)
            torch.cuda.synchronize()
            st = ha.get_state(committed=False)
            pl = {"workload": f"Tet10 block, {ma.n_elems} elements, {ma.n_dofs} DOF, {4 * ma.n_elems} Gauss points, J2 plasticity with linear kinematic "
                              "hardening (BASELINE.json configs[4])", "yielded_fraction": float(np.mean(st[:, :, 12] > 0)),
                  "state_update_ms": ms_state, "gauss_points_per_s": 4 * ma.n_elems / (ms_state * 1e-3),
                  "tangent_assembly_ms": ms_tan, "elements_per_s": ma.n_elems / (ms_tan * 1e-3),
                  "frac_hbm_roofline": ASSEMBLY_BYTES_PER_ELEM * ma.n_elems / (ms_tan * 1e-3) / 1e9 / peak,
                  "assembled_vs_matfree_tangent_rel": float((y1 - y2).abs().max() / y2.abs().max()),
                  "l2": "flushed between repetitions" if flush is not None else "not flushed"}
            return asm, pl
        finally:
            ha.close()

    def linear_static_leg():
        """BASELINE.json configs[0] = examples/linear_static.jl of the reference, the one end-to-end value the reference itself holds
        (max |u| = 2.4052929896922337, :133): committed copy of its mesh (tests/golden/linear_static_smp18.npz), E = 208e3,
        nu = 0.3, body load 1.0 in x, the clamp of :46-72; consistent load on the device, projected CG to 1e-12.  CPU side: the
        oracle's assembly + a sparse direct solve (scipy splu) as the stand-in for the reference's LDLt (src/solvers.jl:205-210)."""
        z = np.load(os.path.join(ROOT, "tests", "golden", "linear_static_smp18.npz"))
        m = mesh.Mesh(10, z["coords"], z["conn"])
        nodes = np.union1d(mesh.nodes_at_plane(m, 1, 50.0, 6.0), mesh.find_nearest_nodes(m, [165.0, 88.0, 10.0], 3))
        fixed = np.sort((3 * (nodes[:, None] - 1) + np.arange(1, 4)[None, :]).ravel())
        ref = 2.4052929896922337
        t0 = time.perf_counter()
        hl = _lib.Handle(10, m.coords, m.conn, device=local_rank)
        try:
            hl.set_material(_lib.MAT_LINEAR_ELASTIC, (208.0e3, 0.30))
            hl.set_dirichlet(fixed)
            hl.set_stream(torch.cuda.current_stream().cuda_stream)
            f = hl.body_load((1.0, 0.0, 0.0))
            u, it, res = hl.cg(f, tol=1e-12, relative=True, max_iter=200000)
            torch.cuda.synchronize()
            t_gpu = time.perf_counter() - t0
        finally:
            hl.close()
        umax = float(np.linalg.norm(np.asarray(u).reshape(-1, 3), axis=1).max())
        out_ = {"workload": f"examples/linear_static.jl (BASELINE.json configs[0]): {m.n_nodes} nodes, {m.n_elems} Tet10, {fixed.size} fixed dofs",
                "max_u_norm": umax, "reference_value": ref, "rel_err": abs(umax / ref - 1.0), "tol": 1.5e-8,
                "ok": bool(abs(umax / ref - 1.0) < 1.5e-8), "cg_iterations": int(it), "final_abs_residual": float(res),
                "gpu_seconds_setup_load_solve": t_gpu,
                "note": "reference value and tolerance: @test isapprox(maximum(values(u_norms)), 2.4052929896922337), examples/linear_static.jl:133"}
        if do_cpu:
            import scipy.sparse as sp
            import scipy.sparse.linalg as spla
            from oracle import oracle as O
            O.set_num_threads(host_threads())
            t0 = time.perf_counter()
            rp, ci, vals, _ = O.assemble_csr(10, m.coords, m.conn, par=(208.0e3, 0.30), symmetrise=True)
            fc = O.body_load(10, m.coords, m.conn, (1.0, 0.0, 0.0))
            K = sp.csr_matrix((vals, ci, rp))
            free = np.setdiff1d(np.arange(m.n_dofs), fixed - 1)
            uc = np.zeros(m.n_dofs)
            uc[free] = spla.splu(K[free][:, free].tocsc()).solve(fc[free])
            out_["cpu_seconds_assemble_direct_solve"] = time.perf_counter() - t0
            out_["cpu_max_u_norm"] = float(np.linalg.norm(uc.reshape(-1, 3), axis=1).max())
            out_["gpu_vs_cpu_field_rel"] = float(np.abs(np.asarray(u) - uc).max() / np.abs(uc).max())
        return out_

    cg_out = asm_out = hex_out = nh_out = pl_out = cg_nh = jac_out = ls_out = None
    if not args.no_extras:
        cg_wl = args.cg
        if cg_wl == "auto":
            cg_wl = "T10" if world == 1 else "same"
        if cg_wl == "same" or cg_wl == args.workload:
            cg_out = cg_solve(pp, workload_label(args.workload, world))
        elif cg_wl != "none":
            pp2, _ = setup_problem(cg_wl)
            cg_out = cg_solve(pp2, workload_label(cg_wl, world))
            if world == 1 and args.neohooke:
                cg_nh = guarded(lambda: neo_hookean_leg(pp2, workload_label(cg_wl, world)))
            pp2.handle.close()
        hx = args.hex8
        if hx == "auto":
            hx = "H12" if WORKLOADS[args.workload][0] == 10 else "none"

        def hex8_leg():
            pp3, u3 = setup_problem(hx)
            x3 = torch.from_numpy(u3).to(dev)
            y3 = torch.empty_like(x3)
            par3 = check_parity(pp3, u3, x3, y3) if (do_cpu and not args.no_parity) else None
            t3, _ = time_matvec(pp3, x3, y3, max(5, args.steps // 2), 3, 0.2)
            ms3 = allmax(float(t3.mean()))
            nd3 = 3 * pp3.n_nodes_global
            peak, _ = peaks()
            res_hex = {"workload": workload_label(hx, world) + " (BASELINE.json configs[2]: 12.5 M DOF per GPU, weak scaling; 99.6 M DOF at 8 GPUs)",
                       "value": nd3 / (ms3 * 1e-3) / 1e9, "unit": "GDOF/s", "ms_per_step": ms3, "n_gpus": world, "dofs": nd3,
                       "roofline_frac": BYTES_PER_DOF[8] * nd3 / world / (ms3 * 1e-3) / 1e9 / peak, "parity": par3,
                       "setup_s": float(pp3.handle.info().setup_seconds), "n_patches": int(pp3.handle.info().n_patches)}
            pp3.handle.close()
            return res_hex
        if hx != "none":
            hex_out = guarded(hex8_leg) if world == 1 else hex8_leg()
        if world == 1 and cg_nh is not None:
            nh_out = cg_nh
        if world == 1 and args.assembly != "none":
            asm_out, pl_out = guarded(lambda: assembly_legs(args.assembly), pair=True)
        if world == 1 and args.linear_static:
            ls_out = guarded(linear_static_leg)
        if world == 1 and args.jacobi and cg_wl not in ("none",):
            # last leg (a failure here cannot take another leg with it): the same solve with the block-Jacobi preconditioner
            def jacobi_leg():
                wl = args.workload if cg_wl == "same" else cg_wl
                pp4, _ = setup_problem(wl)
                try:
                    return cg_solve(pp4, workload_label(wl, world), flags=_lib.JACOBI)
                finally:
                    pp4.handle.close()
            jac_out = guarded(jacobi_leg)

    if rank == 0:
        peak, peak_src = peaks()
        gdofs = total_dofs / (ms * 1e-3) / 1e9
        ach = BYTES_PER_DOF[et] * total_dofs / world / (ms * 1e-3) / 1e9     # per GPU
        out = {
            "metric": METRIC, "value": gdofs, "unit": "GDOF/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {'Tet10' if et == 10 else 'Hex8'} matrix-free K.u, {total_dofs} DOF, {total_elems} elements, "
                                   f"linear elastic E=210e9 nu=0.3, clamp x=0, deterministic scatter",
                       "l2": "flushed between timed steps (512 MiB write)" if flush is not None else "not flushed",
                       "patch_elems": int(info.patch_elems), "n_patches": int(info.n_patches), "affine_elems": int(info.n_affine_elems), "smem_bytes": int(info.smem_bytes), "blocks_per_sm": int(info.blocks_per_sm), "interface_nodes": int(info.n_interface_nodes),
                       "partition": ("z-slabs by contiguous node range, owner-computes + ghost elements; halo via "
                                     + ("ncclSend/ncclRecv" if args.nccl_halo else "peer-memory stores over NVLink (CUDA IPC) issued from inside the patch kernel; "
                                        "patches that read ghost values run last and wait for the neighbours' flags")) if world > 1 else "single GPU",
                       "ms_min": float(times.min()), "ms_max": float(times.max()), "setup_s": float(info.setup_seconds),
                       "launch": "one CUDA graph replay of the K timed steps" if tinfo["use_graph"] else "per-step host launches",
                       "host_enqueue_ms_per_step": None if tinfo["host_s"] is None else tinfo["host_s"] * 1e3,
                       "per_rank_ms_mean_min_median_max": per_rank, "rank0_first_steps_ms": [round(float(v), 4) for v in times[:16]],
                       "host_threads": host_threads(), "clock_ramp": tinfo.get("clock_ramp")},
            "parity": parity,
            "e2e": {"value": total_dofs / e2e_s / 1e9, "unit": "GDOF/s", "h2d_bytes_per_step": 8 * n_local_dofs, "d2h_bytes_per_step": 8 * n_local_dofs,
                    "ms_per_step": e2e_s * 1e3, "checksum_abs_y": checksum},
            "gpu_launches": tinfo["launches"],
            "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                         "peak_source": peak_src, "bytes_per_dof": BYTES_PER_DOF[et],
                         "note": "achieved = algorithmic bytes of one K.u / CUDA-event time of the whole step (patch kernel + interface reduce)"},
            "fp64": {"tflops": FP64_FLOP_PER_ELEM[et] * total_elems / world / (ms * 1e-3) / 1e12,
                     "flop_per_element": FP64_FLOP_PER_ELEM[et],
                     "note": "fp64 instructions of the closed-form element kernel (SASS count, DFMA = 2 flop) per GPU; measured DFMA peak of the "
                             "pipe: 64 lanes/clk/SM = 37.2 TFLOP/s at 1965 MHz (profiles/microbench/fp64_pipe.cu)"},
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
        if jac_out is not None:
            out["cg_time_to_solve_block_jacobi"] = jac_out
        if ls_out is not None:
            out["linear_static"] = ls_out
        if asm_out is not None:
            out["assembly"] = asm_out
        if pl_out is not None:
            out["plasticity"] = pl_out
        if nh_out is not None:
            out["neo_hookean"] = nh_out
        if hex_out is not None:
            out["hex8_weak"] = hex_out
        if world == 1 and do_cpu:
            out["cpu_baseline"] = cpu_baseline(args.workload, threads=host_threads())
        print(json.dumps(out), flush=True)
    h.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
