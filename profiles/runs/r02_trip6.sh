set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t6_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t6_pytest.log
tail -15 gpurun_out/t6_pytest.log
for cfg in "T1 ws " "T10 ws " "H12 ws "; do
  set -- $cfg; wl=$1; tag=$2; shift 2
  timeout 300 python bench.py --workload $wl --no-cpu --steps 30 --warmup 3 "$@" > gpurun_out/t6_bench_${wl}_${tag}.json 2> gpurun_out/t6_bench_${wl}_${tag}.err
  tail -c 400 gpurun_out/t6_bench_${wl}_${tag}.err
done
timeout 120 python profiles/phase_timing.py > gpurun_out/t6_phase.log 2>&1
JFEM_SKIP=4 timeout 120 python profiles/phase_timing.py > gpurun_out/t6_phase_skip4.log 2>&1
cat gpurun_out/t6_phase.log gpurun_out/t6_phase_skip4.log | grep -v Traceback
for f in gpurun_out/t6_bench_*.json; do python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', round(d['value'],2), round(d['ms_per_step']*1e3,1),'us frac', round(d['roofline']['frac'],3), 'patches',d['config']['n_patches'], 'smem',d['config']['smem_bytes'], 'e2e', round(d['e2e']['value'],2))"; done
