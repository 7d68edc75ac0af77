python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k plane 2>&1 | tail -30
python -m pytest tests -m gpu -x -q -k "8192 or long_haul or fft_variants" 2>&1 | tail -3
