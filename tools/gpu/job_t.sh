PYATM_FFT_DIRECT=1 python tools/gpu/fft_variants.py --sizes 2048 --only cold2 cold4 2>&1 | cut -c1-400
