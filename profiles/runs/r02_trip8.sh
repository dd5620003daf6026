cd $GRAFT_REPO_ROOT
( time timeout 800 python bench.py --steps 50 --warmup 5 > gpurun_out/t8_bench_default.json 2> gpurun_out/t8_bench_default.err ) 2> gpurun_out/t8_time.txt
tail -c 600 gpurun_out/t8_bench_default.err; cat gpurun_out/t8_time.txt
python - <<'PY'
import json
d=json.loads(open('gpurun_out/t8_bench_default.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','parity','e2e','cg_time_to_solve','assembly','hex8_weak','fp64','cpu_baseline','clocks'):
    print(k, json.dumps(d.get(k))[:600])
print(d['roofline'])
PY
