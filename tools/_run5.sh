cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > gpurun_out/s8_pytest_all.log
timeout 300 python tools/tune.py --variant canny --shapes 512x224,256x288,128x224,64x224,32x224 --ths 0 > gpurun_out/s8_tune_auto.log 2>&1
timeout 300 python tools/tune.py --variant bpda --shapes 512x224,256x288 --ths 0 >> gpurun_out/s8_tune_auto.log 2>&1
cat gpurun_out/s8_pytest_all.log gpurun_out/s8_tune_auto.log
