mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -s > gpurun_out/r2_gpu_tests_a.log 2>&1; echo "pytest rc $?"; tail -5 gpurun_out/r2_gpu_tests_a.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke_a.log 2>&1; echo "smoke rc $?"; tail -3 gpurun_out/r2_smoke_a.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err; echo "bench rc $?"; tail -c 3000 gpurun_out/r2_bench_a.json; tail -5 gpurun_out/r2_bench_a.err
python bench.py --workload c4 > gpurun_out/r2_c4_n1_a.json 2> gpurun_out/r2_c4_n1_a.err; echo "c4 rc $?"; cat gpurun_out/r2_c4_n1_a.json; tail -5 gpurun_out/r2_c4_n1_a.err
python bench.py --workload c5 > gpurun_out/r2_c5_n1_a.json 2> gpurun_out/r2_c5_n1_a.err; echo "c5 rc $?"; cat gpurun_out/r2_c5_n1_a.json; tail -5 gpurun_out/r2_c5_n1_a.err
