cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2.py -x -q 2>&1 | tail -15 > gpurun_out/r2h_pytest.log
cat gpurun_out/r2h_pytest.log
