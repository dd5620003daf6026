set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/t1_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t1_pytest.log
timeout 120 profiles/microbench/phase1_occupancy > gpurun_out/t1_phase1.log 2>&1
for cfg in "T1 ws1_256 " "T1 ws0_128 --patch 128 --opt warp_specialised=0" "T1 ws0_256 --opt warp_specialised=0" "T1 ws0_512 --patch 512 --opt warp_specialised=0" "T10 ws1_256 " "T10 ws0_512 --patch 512 --opt warp_specialised=0" "T10 ws0_256 --opt warp_specialised=0"; do
  set -- $cfg; wl=$1; tag=$2; shift 2
  timeout 300 python bench.py --workload $wl --no-cpu --steps 30 --warmup 3 "$@" > gpurun_out/t1_bench_${wl}_${tag}.json 2> gpurun_out/t1_bench_${wl}_${tag}.err
done
tail -3 gpurun_out/t1_pytest.log; cat gpurun_out/t1_phase1.log
for f in gpurun_out/t1_bench_*.json; do echo $f; python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['config']['n_patches'], d['config']['interface_nodes'], d['config']['smem_bytes'], d['config']['blocks_per_sm'])"; done
