mkdir -p gpurun_out
python tools/gpu/fft_variants.py --sizes 2048 2>&1 | tee gpurun_out/r2_fft_variants_k.log
python -m pytest tests -m gpu -x -q > gpurun_out/r2_gpu_tests_k.log 2>&1; echo "all tests rc $?"; tail -3 gpurun_out/r2_gpu_tests_k.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_k.json 2> gpurun_out/r2_bench_k.err; echo "bench rc $?"; tail -2 gpurun_out/r2_bench_k.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_k.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e'], d['roofline']['per_kernel_us'], d['roofline']['frac'], d['roofline_screen']['us_per_screen'], d['clocks'])"
python bench.py --workload c4 2>/dev/null | cut -c1-300
python bench.py --workload c5 2>/dev/null | cut -c1-300
