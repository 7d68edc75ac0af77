mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_cols_tma -s 3 -c 1 -f -o gpurun_out/r2_cols python tools/gpu/prof_pass.py cols > gpurun_out/r2_ncu_cols.log 2>&1; tail -2 gpurun_out/r2_ncu_cols.log
ncu --set full --clock-control none --import-source on -k regex:k_rows -s 3 -c 1 -f -o gpurun_out/r2_rows python tools/gpu/prof_pass.py rows > gpurun_out/r2_ncu_rows.log 2>&1; tail -2 gpurun_out/r2_ncu_rows.log
python tools/ncu_summary.py gpurun_out/r2_cols.ncu-rep 30 > gpurun_out/r2_prof_cols_summary.txt 2>&1
python tools/ncu_summary.py gpurun_out/r2_rows.ncu-rep 30 > gpurun_out/r2_prof_rows_summary.txt 2>&1
python tools/gpu/fft_variants.py --sizes 1024 2048 4096 --only default 2>&1 | tee gpurun_out/r2_fft_variants_f.log
python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r2_bench_f.json 2> gpurun_out/r2_bench_f.err; echo "bench rc $?"; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_f.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['per_kernel_us'], d['roofline_screen']['us_per_screen'])"
ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 400 --csv --log-file gpurun_out/r2_launches_f.csv python tools/gpu/prof_pass.py step 2048 8 30 > /dev/null 2>&1; python tools/launch_breakdown.py gpurun_out/r2_launches_f.csv 2>&1 | tail -25
