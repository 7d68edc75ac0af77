mkdir -p gpurun_out
python tools/gpu/fft_variants.py --sizes 2048 4096 8192 --env PYATM_FFT_ROWS_TMA=1 2>&1 | tee gpurun_out/r2_fft_variants_l.log
for c in 32 64; do echo "chunk $c:"; PYATM_SIM_CHUNK=$c python bench.py --steps 20 --warmup 5 --no-cpu 2>/dev/null | python -c "
import json,sys
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); print(round(d['value'],1), round(d['e2e']['value'],1), d['ms_per_step'], d['clocks']['sm_mhz'], d['roofline']['per_kernel_us'], d['roofline']['frac'], d['roofline']['traffic'])"; done
python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k plane 2>&1 | tail -2
