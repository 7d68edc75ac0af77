python -m pytest tests -m gpu -x -q > gpurun_out/r2_gpu_tests_final.log 2>&1; echo "all tests rc $?"; tail -3 gpurun_out/r2_gpu_tests_final.log
python - <<'PY'
import time, numpy as np, torch
import pyatmosphere_b200 as pa
from bench import C3, build_channel
for name, mk, n in (("README QuickChannel 1024^2", lambda: pa.QuickChannel(Cn2=1e-15, length=10000, count_ps=5, beam_w0=0.09, beam_wvl=8.08e-07, aperture_radius=0.12), 2000),
                    ("advanced channel 2048^2", lambda: build_channel(pa, C3), 1000)):
    for rng in ("numpy", "philox"):
        pa.gpu.config.update(use_gpu=True, rng=rng, seed=3, batch=32)
        ch = mk()
        for rep in range(2):
            beam = pa.simulations.BeamResult(ch, max_size=n); pdt = pa.simulations.PDTResult(ch, max_size=n)
            np.random.seed(1); torch.cuda.synchronize(); t0 = time.perf_counter()
            pa.simulations.Simulation([beam, pdt]).run()
            dt = time.perf_counter() - t0
        print(f"{name}, rng={rng}: {n/dt:.0f} realizations/s; sigma_BW {beam.bw[0]:.4f}", flush=True)
PY
