mkdir -p gpurun_out
./tools/micro/tma_swizzle_probe > gpurun_out/r2_tma_swizzle_probe.log 2>&1; echo "probe rc $?"; cat gpurun_out/r2_tma_swizzle_probe.log | head -120
PYATM_COLS_SW=0 python tools/gpu/fft_variants.py --sizes 2048 2>&1 | tee gpurun_out/r2_fft_variants_c.log
