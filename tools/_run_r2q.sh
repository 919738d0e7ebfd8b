cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu -x 2>&1 | tail -4 > gpurun_out/r2q_pytest.log; cat gpurun_out/r2q_pytest.log
timeout 300 python tools/tune_attacks.py 2>&1 | sed -n '/per-call/,$p' > gpurun_out/r2q_named.log; cat gpurun_out/r2q_named.log
timeout 300 python bench.py --no-e2e --no-cpu-baseline --no-extra --no-named-batch 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], {k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()})"
