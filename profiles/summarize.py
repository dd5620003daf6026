"""Turns an ncu report (brought back in gpurun_out/) into the small tracked summaries under profiles/.

    python profiles/summarize.py gpurun_out/prof.ncu-rep r01_patch_kernel_T1 [launch_index]
    python profiles/summarize.py --launches gpurun_out/launches.csv r01_launches_T1
"""
import csv
import json
import os
import subprocess
import sys
from collections import defaultdict

HERE = os.path.dirname(os.path.abspath(__file__))
KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.per_cycle_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__sass_inst_executed_op_shared_ld.sum", "smsp__sass_inst_executed_op_shared_st.sum", "smsp__sass_inst_executed_op_global_ld.sum",
        "smsp__sass_inst_executed_op_global_st.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def launches(path, name):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    d = defaultdict(list)
    for r in rows[1:]:
        try:
            d[r[ik]].append(float(r[iv].replace(",", "")))
        except ValueError:
            pass
    tot = sum(sum(v) for v in d.values())
    out = [{"kernel": k[:120], "launches": len(v), "avg_ns": sum(v) / len(v), "min_ns": min(v), "max_ns": max(v), "share": sum(v) / tot} for k, v in d.items()]
    out.sort(key=lambda x: -x["share"])
    json.dump(out, open(os.path.join(HERE, name + ".json"), "w"), indent=1)
    for o in out:
        print(f'{o["share"]:.3f} {o["avg_ns"] / 1e3:9.2f} us x{o["launches"]:3d}  {o["kernel"][:90]}')


def report(path, name, idx=0):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2 + idx]
    out = {"kernel": vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"}
    for i, k in enumerate(hdr):
        if k in KEYS:
            out[k] = {"value": vals[i], "unit": units[i]}
    stalls = []
    for i, k in enumerate(hdr):
        if "pcsamp_warps_issue_stalled" in k and "not_issued" not in k:
            try:
                stalls.append((float(vals[i]), k.replace("smsp__pcsamp_warps_issue_stalled_", "")))
            except ValueError:
                pass
    stalls.sort(reverse=True)
    out["stall_samples"] = {k: v for v, k in stalls[:10]}
    json.dump(out, open(os.path.join(HERE, name + ".json"), "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    if sys.argv[1] == "--launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        report(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 0)
