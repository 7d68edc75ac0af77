mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_fft_variants.py -m gpu -x -q -k "vacuum or long_haul or split_column or fft" > gpurun_out/r2_fft_tests_b.log 2>&1; echo "fft tests rc $?"; tail -4 gpurun_out/r2_fft_tests_b.log
python tools/gpu/fft_variants.py --sizes 2048 --env PYATM_COLS_SW=0 2>&1 | tee gpurun_out/r2_fft_variants_b.log
python tools/gpu/fft_variants.py --sizes 1024 4096 8192 --only default 2>&1 | tee -a gpurun_out/r2_fft_variants_b.log
timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "test_vacuum_leg_random_field_vs_oracle and 2048 and complex64" > gpurun_out/r2_racecheck_cols_sw.log 2>&1; echo "racecheck rc $?"; tail -6 gpurun_out/r2_racecheck_cols_sw.log
python -m pytest tests -m gpu -x -q > gpurun_out/r2_gpu_tests_b.log 2>&1; echo "all tests rc $?"; tail -4 gpurun_out/r2_gpu_tests_b.log
