set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "test_edge_filter_fwd_bwd and (canny or bpda) and hyst and (8 or 9 or 0)" 2>&1 | tail -3 > gpurun_out/s7_pytest.log
timeout 300 python tools/tune.py --variant canny --shapes 512x224,256x288,128x224,64x224,32x224 --ths 0,28,56,112,224 > gpurun_out/s7_tune_stream.log 2>&1
timeout 300 python tools/tune.py --variant canny --shapes 128x224,64x224 --ths 0 --staging 9 > gpurun_out/s7_tune_tiles.log 2>&1
cat gpurun_out/s7_pytest.log gpurun_out/s7_tune_stream.log gpurun_out/s7_tune_tiles.log
