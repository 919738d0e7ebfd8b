cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu -x 2>&1 | tail -6 > gpurun_out/r2l_pytest.log; cat gpurun_out/r2l_pytest.log
timeout 600 python tools/train_throughput.py --front all --iters 6 > gpurun_out/r2l_train_tiny.jsonl 2> gpurun_out/r2l_train.err; cat gpurun_out/r2l_train_tiny.jsonl
timeout 600 python tools/train_throughput.py --front all --iters 4 --config imagenet_free > gpurun_out/r2l_train_imagenet_free.jsonl 2>> gpurun_out/r2l_train.err; cat gpurun_out/r2l_train_imagenet_free.jsonl
for s in "4096 64" "256 64" "32 224" "128 28"; do timeout 300 python tools/front_end_time.py $s >> gpurun_out/r2l_front_end.jsonl 2>> gpurun_out/r2l_train.err; done; cat gpurun_out/r2l_front_end.jsonl
tail -3 gpurun_out/r2l_train.err
