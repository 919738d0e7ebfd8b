cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2s_bench_n$N.json 2> gpurun_out/r2s_bench_n$N.err
tail -c 200 gpurun_out/r2s_bench_n$N.err
