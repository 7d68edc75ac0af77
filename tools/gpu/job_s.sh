python tools/gpu/fft_variants.py --sizes 2048 4096 2>&1 | tail -2 | cut -c1-600
python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -m gpu -x -q -k "tiny or fused_statistics or time_series" 2>&1 | tail -2
