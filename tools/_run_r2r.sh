cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python bench.py --no-extra --no-named-batch --no-cpu-baseline > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err; tail -c 300 gpurun_out/r2r_bench.err
