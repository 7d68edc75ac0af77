mkdir -p gpurun_out
python -m pytest tests/test_gpu_round2.py -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r2_bench_n.json 2> gpurun_out/r2_bench_n.err; echo "bench rc $?"; tail -2 gpurun_out/r2_bench_n.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_n.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e'], d['roofline']['frac'], d['clocks'])"
