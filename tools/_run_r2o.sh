cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2o_smoke.log 2>&1; tail -2 gpurun_out/r2o_smoke.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/r2o_bench_n4.json 2> gpurun_out/r2o_bench_n4.err
tail -c 300 gpurun_out/r2o_bench_n4.err; python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 4 --steps 2 --warmup 1 > gpurun_out/r2o_bench_ref_n4.json 2>> gpurun_out/r2o_bench_n4.err; tail -c 400 gpurun_out/r2o_bench_ref_n4.json
