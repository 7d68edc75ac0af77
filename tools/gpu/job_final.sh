mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_gpu_tests_final.log 2>&1; echo "all tests rc $?"; tail -3 gpurun_out/r2_gpu_tests_final.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke_final.log 2>&1; echo "smoke rc $?"; tail -2 gpurun_out/r2_smoke_final.log
ncu --set full --clock-control none --import-source on -k regex:k_cols_tma -s 3 -c 1 -f -o gpurun_out/r2_cols python tools/gpu/prof_pass.py cols > gpurun_out/r2_ncu_cols.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_rows -s 3 -c 1 -f -o gpurun_out/r2_rows python tools/gpu/prof_pass.py rows > gpurun_out/r2_ncu_rows.log 2>&1
python tools/ncu_summary.py gpurun_out/r2_cols.ncu-rep 30 > gpurun_out/r2_prof_cols_summary.txt 2>&1; grep -E "dram__bytes|gpu__time" gpurun_out/r2_prof_cols_summary.txt
python tools/ncu_summary.py gpurun_out/r2_rows.ncu-rep 30 > gpurun_out/r2_prof_rows_summary.txt 2>&1; grep -E "dram__bytes|gpu__time" gpurun_out/r2_prof_rows_summary.txt
python tools/gpu/fft_variants.py --sizes 1024 2048 4096 8192 --only default 2>&1 | tee gpurun_out/r2_fft_variants_final.log
python tools/gpu/fft_variants.py --sizes 2048 8192 --only default --dtypes complex128 2>&1 | tee -a gpurun_out/r2_fft_variants_final.log
ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 400 --csv --log-file gpurun_out/r2_launches_final.csv python tools/gpu/prof_pass.py step 2048 8 30 > /dev/null 2>&1; python tools/launch_breakdown.py gpurun_out/r2_launches_final.csv > gpurun_out/r2_step_breakdown_final.txt 2>&1; head -14 gpurun_out/r2_step_breakdown_final.txt
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 200 --csv --log-file gpurun_out/r2_launches_chunk32.csv python tools/gpu/prof_pass.py step 2048 32 20 > /dev/null 2>&1; python tools/launch_breakdown.py gpurun_out/r2_launches_chunk32.csv 2>&1 | sed "s/8 realizations/32 realizations/" > gpurun_out/r2_step_breakdown_chunk32.txt; head -14 gpurun_out/r2_step_breakdown_chunk32.txt
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; echo "bench rc $?"; tail -2 gpurun_out/r2_bench_final.err; cut -c1-2500 gpurun_out/r2_bench_final.json
python bench.py --workload c4 > gpurun_out/r2_c4_final.json 2>/dev/null; cut -c1-200 gpurun_out/r2_c4_final.json
python bench.py --workload c5 > gpurun_out/r2_c5_final.json 2>/dev/null; cut -c1-200 gpurun_out/r2_c5_final.json
