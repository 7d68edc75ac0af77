"""Realization sharding over the GPUs of one box and the reduction of Monte-Carlo statistics.

Independent realizations are the only parallel axis of this path: every batch of realizations is cut into
G contiguous blocks of global indices, rank r takes block r, and the device RNG is keyed by the global index, so the
records do not depend on G.  The data path has no collective; only the per-sample scalar tables (a few KB) are gathered and
the PDT histograms / moment sums all-reduced, over NCCL on GPUs (gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np


def _td():
    import torch.distributed as td
    return td


def world_rank():
    try:
        td = _td()
        if td.is_available() and td.is_initialized():
            return td.get_world_size(), td.get_rank()
    except ImportError:
        pass
    return 1, 0


def shard_indices(first: int, count: int, rank: int, world: int) -> np.ndarray:
    """Global realization indices in [first, first+count) owned by `rank`: contiguous blocks, sizes differing by
    at most one (so that one launch of the device RNG, keyed by consecutive indices, serves a rank)."""
    return np.array_split(np.arange(first, first + count, dtype=np.int64), world)[rank]


def _comm_device():
    import torch
    td = _td()
    if td.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def gather_rows(local: np.ndarray, local_rows: np.ndarray, total_rows: int) -> np.ndarray:
    """Assemble a [total_rows][C] table on every rank from each rank's rows (`local_rows` = positions of the
    rows of `local` in the full table).  Implemented as an all-reduce(sum) of a zero-filled table: rows are
    disjoint, so the sum is exact and every rank gets identical bytes."""
    world, _ = world_rank()
    local = np.asarray(local, dtype=np.float64)
    if world == 1:
        out = np.zeros((total_rows, local.shape[1]), dtype=np.float64)
        out[local_rows] = local
        return out
    import torch
    td = _td()
    full = torch.zeros((total_rows, local.shape[1]), dtype=torch.float64)
    full[torch.as_tensor(local_rows, dtype=torch.long)] = torch.as_tensor(local)
    full = full.to(_comm_device())
    td.all_reduce(full, op=td.ReduceOp.SUM)
    return full.cpu().numpy()


def all_gather_blocks(local, counts) -> np.ndarray:
    """Concatenate every rank's block of rows, rank 0 first: [sum(counts)][C] float64 on the host of EVERY rank.

    `local` is this rank's block -- a torch tensor (CUDA under NCCL: the block goes device -> NVLink -> device and is read
    back once) or a numpy array; `counts[r]` is the number of rows rank r owns.  One collective per call: this is the only
    communication of a Monte-Carlo block (all-gather of the per-sample scalar table, a few hundred KB)."""
    world, rank = world_rank()
    counts = [int(c) for c in counts]
    if world == 1:
        return np.asarray(local.detach().cpu().numpy() if hasattr(local, "detach") else local, dtype=np.float64)
    import torch
    td = _td()
    dev = _comm_device()
    t = local if hasattr(local, "detach") else torch.as_tensor(np.asarray(local, dtype=np.float64))
    t = t.to(device=dev, dtype=torch.float64)
    assert t.shape[0] == counts[rank], (t.shape, counts, rank)
    width, top = t.shape[1], max(counts)
    padded = torch.zeros((top, width), dtype=torch.float64, device=dev)
    padded[:t.shape[0]] = t
    parts = [torch.empty_like(padded) for _ in range(world)]
    td.all_gather(parts, padded)
    host = torch.stack(parts).cpu().numpy()
    return np.concatenate([host[r, :counts[r]] for r in range(world)], axis=0)


def allreduce_sum(values) -> np.ndarray:
    """Sum an integer histogram or a vector of float64 moment sums over all ranks."""
    world, _ = world_rank()
    arr = np.asarray(values)
    if world == 1:
        return arr.copy()
    import torch
    td = _td()
    t = torch.as_tensor(arr).to(_comm_device())
    td.all_reduce(t, op=td.ReduceOp.SUM)
    return t.cpu().numpy()


def reduce_statistics(local_table: np.ndarray, columns: dict, eta_names=(), bins: int = 200):
    """Reduce per-rank sample tables to the statistics BeamResult / PDTResult report, without gathering
    samples: n, sum v, sum v^2 for v in {<x>^2, 4<x^2>, 4<x^2>-4<x>^2} (simulations/beam.py:35-71) and one
    `bins`-bin histogram on [0,1] per aperture (simulations/pdt.py:30-31).  Returns a dict."""
    t = np.asarray(local_table, dtype=np.float64).reshape(-1, max(1, len(columns)))
    out = {}
    if "mean_x" in columns and "mean_x2" in columns:
        bw2 = t[:, columns["mean_x"]] ** 2
        lt2 = 4 * t[:, columns["mean_x2"]]
        st2 = lt2 - 4 * bw2
        sums = np.array([len(t)] + [f(v) for v in (bw2, lt2, st2) for f in (np.sum, lambda a: np.sum(a * a))], dtype=np.float64)
        sums = allreduce_sum(sums)
        n = sums[0]
        for i, name in enumerate(("bw", "lt", "st")):
            s1, s2 = sums[1 + 2 * i], sums[2 + 2 * i]
            mean = s1 / n
            var = max((s2 - n * mean * mean) / (n - 1), 0.0) if n > 1 else float("nan")
            root = np.sqrt(mean)
            out[name] = (float(root), float(np.sqrt(var) / np.sqrt(n) / 2 / root))
        out["count"] = int(n)
    for name in eta_names:
        h = np.histogram(t[:, columns[name]], bins=bins, range=(0, 1))[0].astype(np.int64)
        out[("hist", name)] = allreduce_sum(h)
    return out


class StatsComm:
    """The collective of the path behind the C ABI (pa_comm_* / pa_stats_allreduce, include/pyatm_b200.h): an NCCL communicator
    owned by libpyatm_b200.so.  The 128-byte id is made on rank 0 and reaches the other ranks through whatever side channel
    the host has -- here torch.distributed's object broadcast when a process group is up, or the `unique_id` argument.

        comm = StatsComm()                       # under torchrun, after init_process_group
        comm.allreduce(hist_dev, sums_dev)       # in place, on the current stream
    """

    def __init__(self, rank=None, world=None, unique_id=None, device=None):
        import ctypes as C
        from . import _native as nat
        torch = nat.torch_mod()
        lib = nat.load()
        w, r = world_rank()
        self.rank = r if rank is None else int(rank)
        self.world = w if world is None else int(world)
        self.device = torch.cuda.current_device() if device is None else int(device)
        if unique_id is None:
            buf = (C.c_ubyte * nat.COMM_ID_BYTES)()
            if self.rank == 0:
                nat.check(lib.pa_comm_unique_id(C.cast(buf, C.c_void_p)))
            box = [bytes(buf)]
            if self.world > 1:
                _td().broadcast_object_list(box, src=0)
            unique_id = box[0]
        self.unique_id = bytes(unique_id)
        idbuf = (C.c_ubyte * nat.COMM_ID_BYTES).from_buffer_copy(self.unique_id)
        self.handle = C.c_void_p()
        nat.check(lib.pa_comm_create(C.byref(self.handle), self.device, self.rank, self.world, C.cast(idbuf, C.c_void_p)))
        self.lib = lib

    def allreduce(self, hist_dev=None, sums_dev=None):
        """Sum a uint64/int64 histogram tensor and a float64 vector over the ranks, in place, on the current stream."""
        from . import _native as nat
        nat.check(self.lib.pa_stats_allreduce(self.handle, nat.ptr(hist_dev), 0 if hist_dev is None else hist_dev.numel(),
                                              nat.ptr(sums_dev), 0 if sums_dev is None else sums_dev.numel(), nat.stream_ptr()))

    def close(self):
        if getattr(self, "handle", None):
            self.lib.pa_comm_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:       # noqa: BLE001
            pass
