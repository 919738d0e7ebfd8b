cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "test_edge_filter_fwd_bwd and (canny or bpda) and hyst and (8 or 0)" 2>&1 | tail -3 > gpurun_out/r2f_pytest.log
timeout 300 python tools/tune.py --variant canny --shapes 512x224,256x288,128x224 --ths 0,56 --staging 8 > gpurun_out/r2f_tune_stream.log 2>&1
cat gpurun_out/r2f_pytest.log gpurun_out/r2f_tune_stream.log
