# final validation of the round-2 build: GPU tests, smoke, default bench (all legs), assembly-kernel evidence
cd $GRAFT_REPO_ROOT
T0=$(date +%s)
timeout 90 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/t11_pytest.log 2>&1; echo "pytest rc=$? t=$(( $(date +%s) - T0 ))s" >> gpurun_out/t11_pytest.log
tail -12 gpurun_out/t11_pytest.log
timeout 40 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/t11_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/t11_smoke.log
T1=$(date +%s)
timeout 150 python bench.py --steps 20 --warmup 5 > gpurun_out/t11_bench_default.json 2> gpurun_out/t11_bench_default.err; echo "bench rc=$? t=$(( $(date +%s) - T1 ))s"
tail -c 600 gpurun_out/t11_bench_default.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/t11_bench_default.json').read().strip().splitlines()[-1])
    for k in ('value','ms_per_step','parity','e2e','cg_time_to_solve','neo_hookean','assembly','plasticity','hex8_weak','roofline'):
        print(k, json.dumps(d.get(k))[:900])
except Exception as e: print('bench parse failed', e)
PY
T2=$(date +%s)
timeout 60 python bench.py --steps 3 --warmup 3 --no-cpu --cg none --hex8 none --assembly small --opt assembly_kernel=0 > gpurun_out/t11_bench_asm_old.json 2> gpurun_out/t11_bench_asm_old.err; echo "old-kernel bench rc=$? t=$(( $(date +%s) - T2 ))s"
python -c "
import json
d=json.loads(open('gpurun_out/t11_bench_asm_old.json').read().strip().splitlines()[-1]); print('OLD', json.dumps(d.get('assembly'))[:500]); print('OLD-PL', json.dumps(d.get('plasticity'))[:400])"
T3=$(date +%s)
timeout 60 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"elem_warp|spmv_kernel|expand_pattern|patch_kernel" -c 60 --csv --log-file gpurun_out/r02_launches_assembly.csv python bench.py --steps 3 --warmup 3 --no-cpu --cg none --hex8 none --assembly small > gpurun_out/t11_launch_asm.log 2>&1; echo "ncu launches rc=$? t=$(( $(date +%s) - T3 ))s"
T4=$(date +%s)
timeout 70 ncu --set full --clock-control none --import-source on -k regex:"elem_warp_kernel" -s 2 -c 1 -o gpurun_out/r02_elem_warp python bench.py --steps 3 --warmup 3 --no-cpu --cg none --hex8 none --assembly small > gpurun_out/t11_ncu_full.log 2>&1; echo "ncu full rc=$? t=$(( $(date +%s) - T4 ))s"
echo "total $(( $(date +%s) - T0 ))s"
