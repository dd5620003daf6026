cd $GRAFT_REPO_ROOT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:patch_kernel -s 4 -c 1 -o gpurun_out/r02a_patch_ws_T1 python bench.py --steps 3 --warmup 3 --no-cpu --graph 0 > gpurun_out/r02a_ncu.log 2>&1
tail -3 gpurun_out/r02a_ncu.log
