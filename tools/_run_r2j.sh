cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python tools/_dbg_hfs.py > gpurun_out/r2j_dbg.log 2>&1; cat gpurun_out/r2j_dbg.log
P="python tools/prof_hfs.py 4096 64 8"
$P > gpurun_out/r2j_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'hfs_kernel' -s 1 -c 1 -f -o gpurun_out/r2j_hfs64 $P > gpurun_out/r2j_ncu.log 2>&1
ls -la gpurun_out/r2j_hfs64.ncu-rep
