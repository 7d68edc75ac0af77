"""Worker of test_two_rank_gloo_reduction: each rank owns a block of a synthetic sample table."""
import sys

import numpy as np
import torch.distributed as td

from pyatmosphere_b200 import distributed as dist
from pyatmosphere_b200 import _engine as eng

rank, world = int(sys.argv[1]), int(sys.argv[2])
td.init_process_group("gloo", rank=rank, world_size=world)
assert dist.world_rank() == (world, rank)
rng = np.random.default_rng(123)
cols = eng.table_columns([0.1], [])
total = 37
full = rng.random((total, len(cols)))
full[:, cols["mean_x"]] -= 0.5
mine = dist.shard_indices(0, total, rank, world)
gathered = dist.gather_rows(full[mine], mine, total)
assert np.array_equal(gathered, full)
# one all-gather of unequal per-rank blocks (the end-of-block collective of Simulation.iter_block), numpy and torch inputs
shares = [len(b) for b in np.array_split(np.arange(total), world)]
start = sum(shares[:rank])
block = full[start:start + shares[rank]]
assert np.array_equal(dist.all_gather_blocks(block, shares), full)
import torch
assert np.array_equal(dist.all_gather_blocks(torch.as_tensor(block), shares), full)
# ragged: fewer rows than ranks -- the last rank owns nothing
few = full[:1]
assert np.array_equal(dist.all_gather_blocks(few if rank == 0 else few[:0], [1] + [0] * (world - 1)), few)
stats = dist.reduce_statistics(full[mine], cols, eta_names=[("fixed", 0.1)])
bw2 = full[:, cols["mean_x"]] ** 2
lt2 = 4 * full[:, cols["mean_x2"]]
for name, v in (("bw", bw2), ("lt", lt2), ("st", lt2 - 4 * bw2)):
    m = np.sqrt(v.mean())
    assert np.isclose(stats[name][0], m, rtol=1e-12)
    assert np.isclose(stats[name][1], v.std(ddof=1) / np.sqrt(total) / 2 / m, rtol=1e-9)
assert stats["count"] == total
h = np.histogram(full[:, cols[("fixed", 0.1)]], bins=200, range=(0, 1))[0]
assert np.array_equal(stats[("hist", ("fixed", 0.1))], h)
td.barrier()
td.destroy_process_group()
print("OK", rank)
