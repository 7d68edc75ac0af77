"""torchrun worker: Simulation.run() sharded over the ranks (NCCL) must give every rank the records of the
single-process run (device RNG keyed by the global realization index)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as td

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pyatmosphere_b200 as pa  # noqa: E402

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
td.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = td.get_rank(), td.get_world_size()
pa.gpu.config.update(dtype="complex64", rng="philox", seed=99, batch=3)
ch = pa.QuickChannel(Cn2=1e-15, length=10000, count_ps=3, beam_w0=0.09, beam_wvl=8.08e-07, aperture_radius=0.12,
                     grid_resolution=512, grid_delta=0.002)
beam = pa.simulations.BeamResult(ch, max_size=10)
pdt = pa.simulations.PDTResult(ch, max_size=14)
pa.simulations.Simulation([beam, pdt]).run()
table = np.array([m.data for m in beam.measures])
etas = np.array(pdt.measures[0].data)
assert table.shape == (6, 10) and etas.shape == (14,)
# every rank holds identical records
t = torch.as_tensor(np.concatenate([table.ravel(), etas]), device="cuda")
ref = t.clone()
td.broadcast(ref, src=0)
assert torch.equal(t, ref), "ranks disagree"
# the collective behind the C ABI (pa_comm_* / pa_stats_allreduce): histogram + sums over the ranks == torch's all-reduce
from pyatmosphere_b200.distributed import StatsComm  # noqa: E402
comm = StatsComm()
hist = torch.arange(200, dtype=torch.int64, device="cuda") * (rank + 1)
sums = torch.tensor([1.5, -2.0, 1e-9], dtype=torch.float64, device="cuda") * (rank + 1)
want_h, want_s = hist.clone(), sums.clone()
td.all_reduce(want_h)
td.all_reduce(want_s)
comm.allreduce(hist, sums)
torch.cuda.synchronize()
assert torch.equal(hist, want_h) and torch.allclose(sums, want_s, rtol=1e-15, atol=0), "pa_stats_allreduce differs from torch.distributed"
comm.close()
if rank == 0:
    np.save(os.environ.get("PYATM_NCCL_OUT", "/tmp/nccl_records.npy"), np.concatenate([table.ravel(), etas]))
td.barrier()
td.destroy_process_group()
print("OK", rank, world)
