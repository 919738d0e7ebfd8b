cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r2c_pytest.log
timeout 600 python bench.py > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2c_bench_ref.json 2>> gpurun_out/r2c_bench.err
timeout 300 python tools/tune_attacks.py --skip-small > gpurun_out/r2c_tune_attacks.log 2>&1
timeout 300 python tools/tune.py --graph --shapes 256x64,64x64,1024x64 --ths 0,16,32 > gpurun_out/r2c_tune_small.log 2>&1
tail -3 gpurun_out/r2c_pytest.log; tail -c 600 gpurun_out/r2c_bench.err; cat gpurun_out/r2c_tune_attacks.log gpurun_out/r2c_tune_small.log
