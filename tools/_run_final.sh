cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu -x 2>&1 | tail -5 > gpurun_out/r2u_pytest.log; cat gpurun_out/r2u_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2u_bench_ref.json 2> gpurun_out/r2u_bench.err
(time timeout 900 python bench.py > gpurun_out/r2u_bench.json 2>> gpurun_out/r2u_bench.err) 2>&1 | grep real
tail -c 300 gpurun_out/r2u_bench.err
