cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2n_smoke.log 2>&1; tail -3 gpurun_out/r2n_smoke.log
B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-named-batch --no-extra"
if $B > gpurun_out/r2n_bench_plain.log 2>&1; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2n_launches_raw.csv $B > gpurun_out/r2n_ncu_bench.log 2>&1
  python tools/launches_summary.py gpurun_out/r2n_launches_raw.csv > gpurun_out/r2n_launches.csv; cat gpurun_out/r2n_launches.csv
fi
P="python tools/prof_one.py --shape 4096x64 --variant step125 --iters 2"
$P > gpurun_out/r2n_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'edge_|ew_kernel' -s 3 -c 3 -f -o gpurun_out/r2n_step125 $P > gpurun_out/r2n_ncu.log 2>&1
python tools/ncu_raw_summary.py gpurun_out/r2n_step125.ncu-rep > gpurun_out/r2n_ncu_full_step125.txt
cat > /tmp/l2prof.py <<'PY'
import torch, sys
sys.path.insert(0, "/root/repo")
from edge_enhancement_b200 import functional as F
for shape in ((4096, 3, 64, 64), (512, 3, 224, 224)):
    x = torch.rand(shape, device="cuda"); g = torch.randn_like(x); x0 = torch.rand_like(x)
    for _ in range(3):
        F.pgd_l2_step(x, g, x0, 0.003, 0.047)
torch.cuda.synchronize(); print("ok")
PY
python /tmp/l2prof.py > gpurun_out/r2n_l2_plain.log 2>&1 && ncu --set full --clock-control none -k regex:'pgd_l2' -s 2 -c 1 -f -o gpurun_out/r2n_l2_64 python /tmp/l2prof.py > gpurun_out/r2n_l2_ncu.log 2>&1
python tools/ncu_raw_summary.py gpurun_out/r2n_l2_64.ncu-rep > gpurun_out/r2n_ncu_full_pgd_l2.txt; cat gpurun_out/r2n_ncu_full_pgd_l2.txt
rm -f gpurun_out/r2n_step125.ncu-rep
