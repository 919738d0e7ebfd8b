cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_hfs.py tests/test_gpu_round2.py -x -q 2>&1 | tail -15 > gpurun_out/r2i_pytest.log
cat gpurun_out/r2i_pytest.log
timeout 300 python tools/hfs_time.py > gpurun_out/r2i_hfs_time.log 2>&1; cat gpurun_out/r2i_hfs_time.log
timeout 300 python tools/front_end_time.py 4096 64 > gpurun_out/r2i_front_end.jsonl 2>&1; cat gpurun_out/r2i_front_end.jsonl
