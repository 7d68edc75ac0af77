mkdir -p gpurun_out
cat > /tmp/c128_cols.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import torch
import pyatmosphere_b200 as pa
from pyatmosphere_b200 import _engine as eng, _native as nat
pa.gpu.config.update(use_gpu=True, dtype="complex128")
ctx = eng.grid_context(pa.RectGrid(8192, 0.75e-3))
f = ctx.empty_field(1); f.zero_()
turns = torch.zeros((1, 8192, 8192), dtype=torch.float64, device="cuda")
for _ in range(4):
    nat.check(ctx.lib.pa_fft_pass(ctx.handle, nat.ptr(f), 1, 0, nat.ptr(turns), 1.0e4, 808e-9, nat.stream_ptr()))
torch.cuda.synchronize()
PY
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_col" -c 12 --csv --log-file gpurun_out/r2_launches_8192_c128.csv python /tmp/c128_cols.py > /dev/null 2>&1
grep -E "k_col" gpurun_out/r2_launches_8192_c128.csv | awk -F'","' '{print $5, $NF}' | cut -c1-120 | tail -6
