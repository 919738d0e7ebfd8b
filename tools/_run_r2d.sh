cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu 2>&1 | tail -25 > gpurun_out/r2d_pytest.log
timeout 600 python bench.py --no-extra --no-named-batch --no-cpu-baseline > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err
tail -12 gpurun_out/r2d_pytest.log; tail -c 300 gpurun_out/r2d_bench.err
