# 2-GPU parity on hardware: owned rows of K.u vs the single-GPU result, CG iteration counts (fused peer-memory halo, NCCL halo)
cd $GRAFT_REPO_ROOT
T0=$(date +%s)
timeout 50 python -m pytest -x -q --durations=4 \
  "tests/test_multi_rank.py::test_nccl_two_gpus_matvec_and_cg[10-p2p]" \
  "tests/test_multi_rank.py::test_nccl_two_gpus_matvec_and_cg[8-p2p]" \
  "tests/test_multi_rank.py::test_nccl_two_gpus_matvec_and_cg[10-nccl]" \
  "tests/test_multi_rank.py::test_nccl_two_gpus_matvec_and_cg[10-p2p-unfused]" > gpurun_out/t12_pytest_2gpu.log 2>&1
echo "pytest rc=$? t=$(( $(date +%s) - T0 ))s" >> gpurun_out/t12_pytest_2gpu.log
nvidia-smi -L >> gpurun_out/t12_pytest_2gpu.log
tail -12 gpurun_out/t12_pytest_2gpu.log
