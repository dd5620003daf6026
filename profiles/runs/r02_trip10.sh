# state of HEAD after the container re-creation (earlier gpurun_out/ was lost): tests, default bench, ncu evidence
cd $GRAFT_REPO_ROOT
T0=$(date +%s)
timeout 110 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/t10_pytest.log 2>&1; echo "pytest rc=$? t=$(( $(date +%s) - T0 ))s" >> gpurun_out/t10_pytest.log
tail -14 gpurun_out/t10_pytest.log
T1=$(date +%s)
timeout 150 python bench.py --steps 20 --warmup 5 > gpurun_out/t10_bench_default.json 2> gpurun_out/t10_bench_default.err; echo "bench rc=$? t=$(( $(date +%s) - T1 ))s"
tail -c 400 gpurun_out/t10_bench_default.err
T2=$(date +%s)
timeout 80 ncu --set full --clock-control none --import-source on -k regex:"patch_kernel|iface_reduce" -s 8 -c 2 -o gpurun_out/r02_patch_ws_T1 python bench.py --steps 3 --warmup 3 --no-cpu --no-extras --spinup 0 --graph 0 > gpurun_out/t10_ncu_full.log 2>&1; echo "ncu full rc=$? t=$(( $(date +%s) - T2 ))s"
T3=$(date +%s)
timeout 70 ncu --metrics gpu__time_duration.sum --clock-control none -c 420 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 10 --warmup 3 --no-cpu --spinup 0 --graph 0 --cg same --hex8 none > gpurun_out/t10_launch.log 2>&1; echo "ncu launches rc=$? t=$(( $(date +%s) - T3 ))s"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/t10_bench_default.json').read().strip().splitlines()[-1])
    for k in ('value','ms_per_step','parity','e2e','cg_time_to_solve','assembly','hex8_weak','fp64','cpu_baseline','clocks','roofline'):
        print(k, json.dumps(d.get(k))[:420])
except Exception as e: print('bench parse failed', e)
PY
echo "total $(( $(date +%s) - T0 ))s"
