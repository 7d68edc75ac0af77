python -m pytest tests -m gpu -x -q > gpurun_out/r2_gpu_tests_final.log 2>&1; echo "all tests rc $?"; tail -3 gpurun_out/r2_gpu_tests_final.log
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -2
python bench.py --impl reference --steps 1 --warmup 0 | cut -c1-400
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; echo "bench rc $?"; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_final.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['traffic'], d['roofline_screen']['frac'], d['cpu_baseline']['value'], d['stats_check']['max_rel_diff'])"
