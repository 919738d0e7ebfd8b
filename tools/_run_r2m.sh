cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "test_edge_filter_fwd_bwd and step125" 2>&1 | tail -3 > gpurun_out/r2m_pytest.log; cat gpurun_out/r2m_pytest.log
for st in 0 9; do timeout 300 python tools/tune.py --variant step125 --shapes 512x224,256x288,128x224,64x224,32x224 --ths 0 --staging $st; done > gpurun_out/r2m_tune.log 2>&1; cat gpurun_out/r2m_tune.log
