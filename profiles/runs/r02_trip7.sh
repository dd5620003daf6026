cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t7_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t7_pytest.log
tail -3 gpurun_out/t7_pytest.log
for wl in T1 T10 H12; do
  timeout 300 python bench.py --workload $wl --no-cpu --steps 30 --warmup 3 > gpurun_out/t7_bench_${wl}.json 2> gpurun_out/t7_bench_${wl}.err
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file gpurun_out/t7_launches.csv python bench.py --steps 20 --warmup 3 --no-cpu --spinup 0 --graph 0 > gpurun_out/t7_launch.log 2>&1
for f in gpurun_out/t7_bench_*.json; do python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', round(d['value'],2), round(d['ms_per_step']*1e3,1),'us frac', round(d['roofline']['frac'],3))"; done
python profiles/summarize.py --launches gpurun_out/t7_launches.csv r02_tmp_launches
