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
