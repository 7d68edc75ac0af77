mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_gpu_tests_i.log 2>&1; echo "all tests rc $?"; tail -3 gpurun_out/r2_gpu_tests_i.log
PYATM_TC_PAIR_TMAP=1 python -m pytest tests/test_gpu_screen_tc.py tests/test_gpu_c3_parity.py -m gpu -x -q -k "not single_cta" > gpurun_out/r2_tmap_tests_i.log 2>&1; echo "tmap tests rc $?"; tail -3 gpurun_out/r2_tmap_tests_i.log
for v in 0 1; do PYATM_TC_PAIR_TMAP=$v ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_screen_tc -s 2 -c 3 --csv --log-file gpurun_out/r2_tc_tmap$v.csv python tools/gpu/prof_pass.py screens > /dev/null 2>&1; echo "tmap=$v:"; grep k_screen_tc gpurun_out/r2_tc_tmap$v.csv | awk -F'","' '{print $NF}' | tr -d '"' | tr '\n' ' '; echo; done
python tools/gpu/fft_variants.py --sizes 2048 2>&1 | tee gpurun_out/r2_fft_variants_i.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 50 python -m pytest tests/test_gpu_screen_tc.py -m gpu -x -q -k "tc_screen_vs_oracle_256" > gpurun_out/r2_racecheck_tc_pair.log 2>&1; echo "racecheck rc $?"; python tools/check_racecheck.py gpurun_out/r2_racecheck_tc_pair.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_i.json 2> gpurun_out/r2_bench_i.err; echo "bench rc $?"; tail -2 gpurun_out/r2_bench_i.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_i.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['per_kernel_us'], d['roofline']['frac'], d['roofline_screen']['us_per_screen'], d['cpu_baseline']['value'], d['stats_check']['max_rel_diff'])"
