cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2g_bench_n2.json 2> gpurun_out/r2g_bench_n2.err
tail -c 400 gpurun_out/r2g_bench_n2.err
timeout 600 python -m pytest tests/test_gpu_properties.py -q -m gpu -k "second_device or DataParallel or data_parallel" 2>&1 | tail -4
nvidia-smi topo -m > gpurun_out/r2g_topo.txt 2>&1; nproc >> gpurun_out/r2g_topo.txt; numactl -H >> gpurun_out/r2g_topo.txt 2>&1; lscpu | head -20 >> gpurun_out/r2g_topo.txt
