mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_fft_variants.py tests/test_gpu_screen_generators.py -m gpu -x -q -k "vacuum or long_haul or split_column or fft or Fft or FFT" > gpurun_out/r2_fft_tests_d.log 2>&1; echo "fft tests rc $?"; tail -4 gpurun_out/r2_fft_tests_d.log
python tools/gpu/fft_variants.py --sizes 2048 1024 2>&1 | tee gpurun_out/r2_fft_variants_d.log
python tools/gpu/fft_variants.py --sizes 4096 8192 --only default 2>&1 | tee -a gpurun_out/r2_fft_variants_d.log
python tools/gpu/fft_variants.py --sizes 1024 2048 4096 8192 --only default --dtypes complex128 2>&1 | tee -a gpurun_out/r2_fft_variants_d.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "test_vacuum_leg_random_field_vs_oracle and (2048 or 1024 or 256) and complex64" > gpurun_out/r2_racecheck_fft.log 2>&1; echo "racecheck rc $?"; tail -6 gpurun_out/r2_racecheck_fft.log
python -m pytest tests -m gpu -x -q > gpurun_out/r2_gpu_tests_d.log 2>&1; echo "all tests rc $?"; tail -4 gpurun_out/r2_gpu_tests_d.log
