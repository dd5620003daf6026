"""bench.py contract, CPU side: the reference arm runs without a GPU and prints one JSON line with the agreed keys."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_contract_line():
    env = dict(os.environ, JFEM_BENCH_CPU_SECONDS="0.5", JFEM_BENCH_CPU_SMALL="1", OMP_NUM_THREADS="1")   # torchrun exports OMP_NUM_THREADS=1
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "3"],
                       capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["metric"] == "tet10_elasticity_matvec_gdofs" and d["unit"] == "GDOF/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["gpu_launches"] == 0
    assert d["config"]["workload"].startswith("T1: Tet10 matrix-free K.u, 1075275 DOF, 255552 elements")
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == d["value"] and "sample" in cb
    # the arm must not inherit torchrun's OMP_NUM_THREADS=1: it sets the thread count itself and reports it
    assert cb["cores"] == len(os.sched_getaffinity(0)) == d["config"]["omp_threads"]
    assert d["e2e"] == {"value": d["value"], "unit": "GDOF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True, text=True,
                       timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_bench_parity_check_rows_on_a_partitioned_mesh():
    """bench.parity_check is the in-run correctness evidence of the multi-GPU runs: every rank compares its owned rows with the
    oracle on its local mesh (owned + ghost elements).  Replayed here without a GPU: the 'GPU result' is the oracle's K.u of
    the WHOLE mesh restricted to the rank, so the check must pass on both the all-rows and the sampled branch (which must contain
    interface rows), and must fail when an interface row is corrupted."""
    import types
    import numpy as np
    sys.path.insert(0, ROOT)
    import bench
    from juliafem.jl_b200 import mesh as M
    from oracle import oracle as O
    dims, box = (3, 2, 6), (1.5, 1.0, 3.0)
    full = M.tet10_kuhn(*dims, *box)
    fixed = M.clamp_dofs(full)
    gd = np.arange(full.n_dofs)
    mask = np.zeros(full.n_dofs, dtype=bool)
    mask[fixed - 1] = True
    u = M.hashed_vector(gd, mask)
    y = O.matfree(10, full.coords, full.conn, u, par=bench.MAT, fixed_dofs=fixed)
    for rank in range(2):
        w, off, nn, ne = M.lattice_window(10, dims, box, 2, rank)
        p = M.partition_mesh(w, 2, rank, off, nn)
        ldofs = (3 * (p.local_nodes[:, None] - 1) + np.arange(3)[None, :]).ravel()
        fixed_local = np.nonzero(mask[ldofs])[0] + 1
        pp = types.SimpleNamespace(n_owned=p.n_owned, conn_local=p.conn_local, elem_type=10, rank=rank, local_nodes=p.local_nodes,
                                   coords_local=np.ascontiguousarray(w.coords[p.local_nodes - 1 - off]), fixed_local=fixed_local)
        y_local, u_local = y[ldofs].copy(), u[ldofs]
        err, n, _ = bench.parity_check(pp, y_local, u_local, 2)
        assert n == 3 * p.n_owned and err < 1e-13
        err, n, _ = bench.parity_check(pp, y_local, u_local, 2, full_limit=10, sample_nodes=60)
        assert 0 < n <= 3 * 60 and err < 1e-13
        # corrupt the row of an owned node that sits next to the partition interface: the sampled check must see it
        ghosty = (p.conn_local > p.n_owned).any(axis=1)
        near = np.unique(p.conn_local[ghosty])
        near = near[near <= p.n_owned]
        free = [k for k in near if not mask[ldofs[3 * (k - 1)]]]
        y_bad = y_local.copy()
        y_bad[3 * (free[0] - 1)] *= 1.0 + 1e-6
        err_all, _, _ = bench.parity_check(pp, y_bad, u_local, 2)
        assert err_all > 1e-12


def test_clock_ramp_guard_logic():
    """bench.wait_for_clocks keeps the GPUs busy until every rank reports >= 90 % of its maximum SM clock (collective decision,
    same step count on every rank; the clock is read while a chunk is in flight) and gives up after max_seconds; without
    NVML it runs its single chunk and stops."""
    sys.path.insert(0, ROOT)
    import bench
    ident = lambda v: float(v)
    steps = []
    # already at full clock: one chunk (extra warm-up), then stop
    r = bench.wait_for_clocks(lambda: steps.append(1), lambda: None, lambda: (1965.0, 1965.0), ident, ident, chunk=10)
    assert r["rounds"] == 1 and r["all_ranks_ramped"] and len(steps) == 10
    # ramps up while we keep it busy
    steps.clear()
    clocks = iter([(120.0, 1965.0), (900.0, 1965.0), (1965.0, 1965.0)])
    r = bench.wait_for_clocks(lambda: steps.append(1), lambda: None, lambda: next(clocks), ident, ident, chunk=10)
    assert r["rounds"] == 3 and len(steps) == 30 and r["all_ranks_ramped"] and r["sm_mhz_under_load"] == 1965.0
    # another rank is still slow (allmin says no) although this one is fine: keeps going until the time limit
    steps.clear()
    r = bench.wait_for_clocks(lambda: steps.append(1), lambda: None, lambda: (1965.0, 1965.0), lambda v: 0.0, ident, chunk=5, max_seconds=0.05)
    assert not r["all_ranks_ramped"] and r["rounds"] >= 1 and len(steps) == 5 * r["rounds"]
    # no NVML: nothing to wait for
    def broken():
        raise RuntimeError("NVML unavailable")
    steps.clear()
    r = bench.wait_for_clocks(lambda: steps.append(1), lambda: None, broken, ident, ident, chunk=7)
    assert r["rounds"] == 1 and len(steps) == 7 and r["sm_mhz_under_load"] is None and r["all_ranks_ramped"]
