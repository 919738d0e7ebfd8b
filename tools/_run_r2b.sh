cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_golden.py tests/test_gpu_attacks_api.py tests/test_gpu_hfs.py tests/test_gpu_add_square.py -x -q 2>&1 | tail -25 > gpurun_out/r2b_pytest.log
timeout 600 python tools/tune_attacks.py > gpurun_out/r2b_tune_attacks.log 2>&1
cat gpurun_out/r2b_pytest.log gpurun_out/r2b_tune_attacks.log
