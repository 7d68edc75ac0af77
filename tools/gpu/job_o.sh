python tools/gpu/fft_variants.py --sizes 1024 2048 4096 8192 2>&1 | tee gpurun_out/r2_fft_variants_o.log
PYATM_LIB=$PWD/pyatmosphere_b200/variants/libpyatm_latestore.so python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "vacuum or long_haul or split_column" 2>&1 | tail -2
