cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
P="python tools/prof_one.py --shape 512x224 --variant canny --iters 2 --staging 8"
$P > gpurun_out/r2e_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'edge_canny_stream' -s 2 -c 2 -f -o gpurun_out/r2e_stream $P > gpurun_out/r2e_ncu.log 2>&1
P2="python tools/prof_one.py --shape 512x224 --variant step125 --iters 2"
$P2 > gpurun_out/r2e_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'edge_' -s 2 -c 2 -f -o gpurun_out/r2e_step125_224 $P2 > gpurun_out/r2e_ncu2.log 2>&1
ls -la gpurun_out/*.ncu-rep
timeout 300 python tools/tune.py --variant canny --shapes 512x224,256x288 --ths 0,56 --staging 8 > gpurun_out/r2e_tune_stream.log 2>&1
timeout 300 python tools/tune.py --variant step125 --shapes 512x224,256x288 --ths 0 >> gpurun_out/r2e_tune_stream.log 2>&1
cat gpurun_out/r2e_tune_stream.log
