cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "test_edge_filter_fwd_bwd and hyst and (8 or 0)" 2>&1 | tail -5 > gpurun_out/r2p_pytest.log; cat gpurun_out/r2p_pytest.log
for lib in "" $GRAFT_REPO_ROOT/tools/ab/lib_old.so; do echo "== lib=$lib"; EDGE_B200_LIB=$lib timeout 300 python tools/tune.py --variant canny --shapes 512x224,256x288,128x224 --ths 0 --staging 8; EDGE_B200_LIB=$lib timeout 300 python tools/tune.py --variant step125 --shapes 512x224 --ths 0 --staging 8; done > gpurun_out/r2p_tune.log 2>&1; cat gpurun_out/r2p_tune.log
