#!/bin/bash
# Run on the GPU box (gpurun):  bash tools/capture_profiles.sh <tag> [variants...]
# 1. launch list of the bench step (gpu__time_duration.sum), aggregated per kernel
# 2. ncu --set full of the fused kernels per variant (tools/prof_one.py, 4096x3x64x64), raw + SASS summaries
# Every ncu command runs only after the same command has exited 0 without ncu.  Text summaries land in gpurun_out/.
set -u
tag=${1:-r1x}; shift || true
variants=${@:-step125 canny}
out=gpurun_out
mkdir -p $out
B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
if $B > $out/${tag}_bench_plain.log 2>&1; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_raw.csv $B > $out/${tag}_ncu_bench.log 2>&1
  python tools/launches_summary.py $out/${tag}_launches_raw.csv > $out/${tag}_launches.csv
fi
shape=${EE_PROF_SHAPE:-4096x64}
stag=${EE_PROF_TAG:-}
for v in $variants; do
  P="python tools/prof_one.py --shape $shape --variant $v --iters 2"
  if $P > $out/${tag}${stag}_prof_${v}_plain.log 2>&1; then
    ncu --set full --clock-control none --import-source on -k regex:'edge_|ew_kernel' -s 3 -c 3 -f -o $out/${tag}${stag}_${v} $P > $out/${tag}${stag}_ncu_${v}.log 2>&1
    python tools/ncu_raw_summary.py $out/${tag}${stag}_${v}.ncu-rep > $out/${tag}${stag}_ncu_full_${v}.txt
    ncu -i $out/${tag}${stag}_${v}.ncu-rep --page source --csv --print-source sass > $out/${tag}${stag}_sass_${v}.csv 2>/dev/null
    python tools/ncu_sass_summary.py $out/${tag}${stag}_sass_${v}.csv > $out/${tag}${stag}_sass_mix_${v}.txt
    python tools/ncu_hot_lines.py $out/${tag}${stag}_sass_${v}.csv > $out/${tag}${stag}_hot_${v}.txt 2>/dev/null
    rm -f $out/${tag}${stag}_${v}.ncu-rep $out/${tag}${stag}_sass_${v}.csv          # gpurun_out/ is capped at 64 MiB
  fi
done
ls -la $out | tail -20
