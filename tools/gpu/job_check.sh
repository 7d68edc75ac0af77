mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_gpu_tests_final.log 2>&1; echo "all tests rc $?"; tail -3 gpurun_out/r2_gpu_tests_final.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke_final.log 2>&1; echo "smoke rc $?"; tail -2 gpurun_out/r2_smoke_final.log
python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; echo "bench rc $?"; tail -2 gpurun_out/r2_bench_final.err; cut -c1-400 gpurun_out/r2_bench_final.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference.json 2>/dev/null; echo "reference arm rc $?"; cut -c1-400 gpurun_out/r2_bench_reference.json
