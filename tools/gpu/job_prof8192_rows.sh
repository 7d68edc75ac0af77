mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_rows -s 2 -c 1 -f -o gpurun_out/r2_rows8192 python tools/gpu/prof_pass.py rows 8192 1 4 > gpurun_out/r2_ncu_rows8192.log 2>&1
python tools/ncu_summary.py gpurun_out/r2_rows8192.ncu-rep 30 > gpurun_out/r2_prof_rows8192_summary.txt 2>&1; head -75 gpurun_out/r2_prof_rows8192_summary.txt
