mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_cols_tma -s 2 -c 1 -f -o gpurun_out/r2_cols8192_inner python tools/gpu/prof_pass.py cols 8192 1 4 > gpurun_out/r2_ncu_cols8192.log 2>&1
python tools/ncu_summary.py gpurun_out/r2_cols8192_inner.ncu-rep 30 > gpurun_out/r2_prof_cols8192_inner_summary.txt 2>&1; head -60 gpurun_out/r2_prof_cols8192_inner_summary.txt
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_col|k_rows" -c 24 --csv --log-file gpurun_out/r2_launches_8192.csv python tools/gpu/prof_pass.py cols 8192 1 6 > /dev/null 2>&1; grep -E "k_col|k_rows" gpurun_out/r2_launches_8192.csv | awk -F'","' '{print $5, $NF}' | cut -c1-120 | tail -9
