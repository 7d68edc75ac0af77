python -m pytest tests -m gpu -x -q > gpurun_out/r2_gpu_tests_final.log 2>&1; echo "all tests rc $?"; tail -3 gpurun_out/r2_gpu_tests_final.log
python - <<'PY'
import time, numpy as np, torch
import pyatmosphere_b200 as pa
for rng in ("numpy", "philox"):
    pa.gpu.config.update(use_gpu=True, rng=rng, seed=3, batch=64)
    ch = pa.QuickChannel(Cn2=1e-15, length=10000, count_ps=5, beam_w0=0.09, beam_wvl=8.08e-07, aperture_radius=0.12)
    n = 2000
    for rep in range(2):
        beam = pa.simulations.BeamResult(ch, max_size=n); pdt = pa.simulations.PDTResult(ch, max_size=n)
        np.random.seed(1); torch.cuda.synchronize(); t0 = time.perf_counter()
        pa.simulations.Simulation([beam, pdt]).run()
        dt = time.perf_counter() - t0
    print(f"README QuickChannel Monte Carlo, rng={rng}: {n/dt:.0f} realizations/s; sigma_BW {beam.bw[0]:.4f}")
PY
