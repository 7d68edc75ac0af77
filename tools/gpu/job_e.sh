mkdir -p gpurun_out
python tools/gpu/fft_variants.py --sizes 2048 4096 8192 2>&1 | tee gpurun_out/r2_fft_variants_e.log
python -m pytest tests -m gpu -x -q > gpurun_out/r2_gpu_tests_e.log 2>&1; echo "all tests rc $?"; tail -4 gpurun_out/r2_gpu_tests_e.log
