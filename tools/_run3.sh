set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "test_edge_filter_fwd_bwd and (canny or bpda) and hyst and (8 or 9 or 0)" 2>&1 | tail -8 > gpurun_out/s6_pytest.log
timeout 300 python tools/tune.py --variant canny --shapes 512x224,256x288,32x224 --ths 0,28,56,112,224 > gpurun_out/s6_tune_stream.log 2>&1
P="python tools/prof_one.py --shape 512x224 --variant canny --iters 2"
$P > gpurun_out/s6_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'edge_canny_stream' -s 2 -c 2 -f -o gpurun_out/s6_stream $P > gpurun_out/s6_ncu.log 2>&1
cat gpurun_out/s6_pytest.log gpurun_out/s6_tune_stream.log
