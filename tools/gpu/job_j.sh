mkdir -p gpurun_out
for c in 4 8 16 32; do echo "chunk $c:"; PYATM_SIM_CHUNK=$c python bench.py --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); print(round(d['value'],1), round(d['e2e']['value'],1), d['ms_per_step'], d['clocks']['sm_mhz'], d['roofline']['per_kernel_us'])"; done
ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 400 --csv --log-file gpurun_out/r2_launches_j.csv python tools/gpu/prof_pass.py step 2048 8 30 > /dev/null 2>&1; python tools/launch_breakdown.py gpurun_out/r2_launches_j.csv 2>&1 | tee gpurun_out/r2_step_breakdown_j.txt | head -8
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py tests/test_gpu_c3_parity.py tests/test_gpu_round2.py -m gpu -x -q -k "(vacuum_leg and 2048 and complex64) or simulate_batch_at_config3 or block_route or stats_allreduce or fused_statistics" > gpurun_out/r2_memcheck_j.log 2>&1; echo "memcheck rc $?"; tail -4 gpurun_out/r2_memcheck_j.log
python - <<'PY'
import time, numpy as np, torch
import pyatmosphere_b200 as pa
pa.gpu.config.update(use_gpu=True, rng="philox", seed=3, batch=64)
ch = pa.QuickChannel(Cn2=1e-15, length=10000, count_ps=5, beam_w0=0.09, beam_wvl=8.08e-07, aperture_radius=0.12)
for n in (2000, 20000):
    beam = pa.simulations.BeamResult(ch, max_size=n); pdt = pa.simulations.PDTResult(ch, max_size=n)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    pa.simulations.Simulation([beam, pdt]).run()
    dt = time.perf_counter() - t0
    print(f"C1 QuickChannel 1024^2 x5 screens: {n} realizations through Simulation.run in {dt:.3f} s = {n/dt:.0f}/s; bw {beam.bw}")
PY
