cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t9_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t9_pytest.log
tail -40 gpurun_out/t9_pytest.log
