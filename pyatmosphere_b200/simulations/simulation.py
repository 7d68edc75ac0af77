"""Monte-Carlo driver.  Mirror of /root/reference/pyatmosphere/simulations/simulation.py:7-152.

Two execution routes give the same records:
  * the generic route (`iter`) follows the reference step by step -- channel.generator per realization, every
    Measure's operations applied to a copy of the output -- and serves arbitrary user operations;
  * the batched route is taken by `run` when every Measure is one of the known reductions (BeamResult / PDTResult /
    TrackedPDTResult records): `gpu.config['batch']` realizations are propagated by one fused call and reduced by one
    sweep.  With `gpu.config['rng'] == 'numpy'` (`iter_batch`) the spectra are drawn from numpy's global RNG in the
    reference's order, so the records equal the generic route's for the same seed.  With 'philox' (`iter_block`) they
    are drawn on the device, keyed by the global realization index: a whole block of realizations -- everything up to
    the next plot / save step, or to the end -- is cut into one contiguous share per rank of the torch.distributed
    group, each rank enqueues its share batch after batch with no host round trip, and the block ends with ONE
    all-gather of the per-sample table (see ..distributed), so every rank appends identical records in index order.
"""
from __future__ import annotations

from typing import Sequence

import numpy as np

from .. import _engine as eng
from .. import _native as nat
from .. import distributed as dist
from .. import gpu
from .measure import Measure
from .result import Result


class Simulation:
    def __init__(self, results_list: Sequence[Result] = None, measures_list: Sequence[Measure] = None):
        self.measures = {}
        self.results_list = results_list
        self.realizations_done = 0          # index of the next realization (device-RNG counter)
        for m in measures_list or []:
            self.add_measures(m)
        for result in results_list or []:
            for m in result.measures:
                self.add_measures(m)
        self._resume_counter()

    def _resume_counter(self):
        """Records loaded from a CSV checkpoint (Result.load_output) were drawn with device-RNG indices 0..L-1: continue
        after them, otherwise a resumed 'philox' run would replay the same realizations and store duplicate samples."""
        done = max((len(m) for m in self.flattened_measures()), default=0)
        self.realizations_done = max(self.realizations_done, done)

    # ---- measure tree: channel -> time -> measure_type -> operations -> [Measure] (simulation.py:19-30) ------
    def add_measures(self, m):
        node = self.measures.setdefault(m.channel, {}).setdefault(m.time, {}).setdefault(m.measure_type, {})
        node.setdefault(m.operations, []).append(m)

    def flattened_measures(self, measures=None):
        measures = measures if measures is not None else self.measures
        if isinstance(measures, dict):
            for values in measures.values():
                yield from self.flattened_measures(values)
        else:
            yield from measures

    def is_measures_done(self, measures=None):
        return all(m.is_done for m in self.flattened_measures(measures))

    # ---- generic route (simulation.py:32-114) ---------------------------------------------------------------
    def init_measures_iteration_data(self):
        for m in self.flattened_measures():
            legs = len(m.channel.path.positions) + 1 if m.measure_type == "propagation" else None
            if m.time:
                m.iteration_data = [[None] * legs for _ in m.time] if legs else [None for _ in m.time]
            else:
                m.iteration_data = [None] * legs if legs else None

    def process_operations(self, output, operations_measures, time_id, propagation_id=None):
        for operations, group in operations_measures.items():
            if self.is_measures_done(group):
                continue
            value = output.copy()
            for op in operations:
                value = op(group[0].channel, output=value)
            for m in group:
                if m.is_done:
                    continue
                if m.time is not None:
                    if propagation_id is not None:
                        m.iteration_data[time_id][propagation_id] = value
                    else:
                        m.iteration_data[time_id] = value
                elif propagation_id is not None:
                    m.iteration_data[propagation_id] = value
                else:
                    m.iteration_data = value

    def iter(self):
        self.init_measures_iteration_data()
        for channel, by_time in self.measures.items():
            for ps in channel.path.phase_screens:
                ps.cache_clear()
            for time, by_type in by_time.items():
                for time_id, t in enumerate(time or [None]):
                    steps = channel.generator(pupil=False, shift=(0, t or 0), store_output=True, wind=True)
                    for leg_id, (field, screen) in enumerate(steps):
                        self.process_operations(field, by_type.get("propagation", {}), time_id, leg_id)
                        if leg_id == 0:
                            self.process_operations(screen, by_type.get("phase_screen", {}), time_id)
                    self.process_operations(channel.output, by_type.get("atmosphere", {}), time_id)
                    self.process_operations(channel.output, by_type.get("propagation", {}), time_id, -1)
                    if channel.pupil:
                        self.process_operations(channel.pupil.output(channel.output), by_type.get("pupil", {}), time_id)
        for m in self.flattened_measures():
            if not m.is_done:
                m.data.append(m.iteration_data)
        self.realizations_done += 1

    # ---- batched route ------------------------------------------------------------------------------------------
    def batchable(self):
        if len(self.measures) != 1:
            return False
        channel = next(iter(self.measures))
        ms = list(self.flattened_measures())
        path = channel.path
        return (all(m.fast_key is not None and m.time is None and m.measure_type == "atmosphere" for m in ms)
                and hasattr(path, "_fusable") and path._fusable() and hasattr(channel.source, "w0"))

    def remaining(self):
        """Realizations still needed (None = unbounded, as in the reference when a max_size is None)."""
        need = 0
        for m in self.flattened_measures():
            if m.max_size is None:
                return None
            need = max(need, m.max_size - len(m))
        return need

    def iter_batch(self, count):
        """`count` realizations in one fused call; appends to every Measure that is not done."""
        channel = next(iter(self.measures))
        ms = list(self.flattened_measures())
        pupils_fixed = sorted({m.fast_key[1] for m in ms if m.fast_key[0] == "pupil_eta" and m.fast_key[2] == "fixed"})
        pupils_tracked = sorted({m.fast_key[1] for m in ms if m.fast_key[0] == "pupil_eta" and m.fast_key[2] == "tracked"})
        world, rank = dist.world_rank()
        first = self.realizations_done
        mine = dist.shard_indices(first, count, rank, world)          # global realization indices of this rank
        table = eng.simulate_realizations(channel, first, count, mine, pupils_fixed, pupils_tracked)
        table = dist.gather_rows(table, mine - first, count)          # [count][columns] on every rank
        cols = eng.table_columns(pupils_fixed, pupils_tracked)
        for r in range(count):
            for m in ms:
                if m.is_done:
                    continue
                key = m.fast_key
                name = key[1] if key[0] == "moment" else (key[2], key[1])
                m.data.append(float(table[r, cols[name]]))
        self.realizations_done += count

    def _fast_keys(self):
        ms = list(self.flattened_measures())
        fixed = sorted({m.fast_key[1] for m in ms if m.fast_key[0] == "pupil_eta" and m.fast_key[2] == "fixed"})
        tracked = sorted({m.fast_key[1] for m in ms if m.fast_key[0] == "pupil_eta" and m.fast_key[2] == "tracked"})
        return ms, fixed, tracked

    def _append_rows(self, ms, table, cols, count):
        for m in ms:
            if m.is_done:
                continue
            key = m.fast_key
            name = key[1] if key[0] == "moment" else (key[2], key[1])
            take = count if m.max_size is None else min(count, m.max_size - len(m))
            m.data.extend(table[:take, cols[name]].tolist())

    def pipelined_host_batches(self, plot_step, save_step):
        """numpy-RNG mode with statistics-only records: batches of gpu.config['batch'] realizations per rank, ONE in flight --
        the spectra of batch k+1 are drawn (numpy's legacy generator, the bottleneck of this mode) while the GPU runs batch k.
        The draws happen in the reference's order whatever the pipelining, and exactly as many as the run needs."""
        channel = next(iter(self.measures))
        ms, fixed, _ = self._fast_keys()
        world, rank = dist.world_rank()
        iteration, pending = 0, None
        while True:
            left = self.remaining()
            ahead = pending.count if pending is not None else 0
            nxt = None
            if left is None or left - ahead > 0:
                count = gpu.config["batch"] * world
                count = count if left is None else min(count, left - ahead)
                for step in (plot_step, save_step):      # a batch never steps across a multiple of the plot / save step
                    if step:
                        count = min(count, step - (iteration + ahead) % step)
                first = self.realizations_done + ahead
                mine = dist.shard_indices(first, count, rank, world)
                nxt = eng.HostBatch(channel, first, count, mine, fixed)
            if pending is not None:
                table = dist.gather_rows(pending.result(), pending.mine - pending.first, pending.count)
                self._append_rows(ms, table, pending.cols, pending.count)
                self.realizations_done += pending.count
                iteration += pending.count
                self.process_output(iteration, plot_step=plot_step, save_step=save_step)
            pending = nxt
            if pending is None:
                return

    def iter_block(self, count):
        """`count` device-RNG realizations: this rank's contiguous share runs with nothing but kernel launches in between
        (engine BlockRunner), then one all-gather hands every rank the whole table."""
        channel = next(iter(self.measures))
        ms = list(self.flattened_measures())
        pupils_fixed = sorted({m.fast_key[1] for m in ms if m.fast_key[0] == "pupil_eta" and m.fast_key[2] == "fixed"})
        pupils_tracked = sorted({m.fast_key[1] for m in ms if m.fast_key[0] == "pupil_eta" and m.fast_key[2] == "tracked"})
        key = (id(channel), tuple(pupils_fixed), tuple(pupils_tracked), gpu.config["dtype"], gpu.config["screen_method"],
               gpu.config["theta_cut"])
        if getattr(self, "_runner_key", None) != key:
            self._runner, self._runner_key = eng.BlockRunner(channel, pupils_fixed, pupils_tracked), key
        world, rank = dist.world_rank()
        first = self.realizations_done
        shares = [len(b) for b in np.array_split(np.arange(count), world)]
        mine_first = first + sum(shares[:rank])
        local = self._runner.run(mine_first, shares[rank])
        self.last_local_table, self.last_columns = local, self._runner.cols        # this rank's share (device), for reductions
        table = dist.all_gather_blocks(local, shares)
        cols = self._runner.cols
        for m in ms:
            if m.is_done:
                continue
            key = m.fast_key
            name = key[1] if key[0] == "moment" else (key[2], key[1])
            take = count if m.max_size is None else min(count, m.max_size - len(m))
            m.data.extend(table[:take, cols[name]].tolist())
        self.realizations_done += count

    # ---- driver (simulation.py:127-152) ----------------------------------------------------------------------
    def run(self, *args, plot_step: int = None, save_step: int = None, **kwargs):
        gpu.require_gpu()
        self._resume_counter()
        try:
            iteration = 0
            use_batch = self.batchable() and gpu.config["batch"] > 1
            use_block = self.batchable() and gpu.config["rng"] != "numpy"
            if use_batch and not use_block:
                _, fixed, tracked = self._fast_keys()
                if not tracked and len(fixed) <= nat.MAX_PUPILS:
                    self.pipelined_host_batches(plot_step, save_step)
            while not self.is_measures_done():
                if use_block:
                    left = self.remaining()
                    count = left if left is not None else 64 * gpu.config["batch"] * dist.world_rank()[0]
                    for step in (plot_step, save_step):      # a block never steps across a multiple of the plot / save step
                        if step:
                            count = min(count, step - iteration % step)
                    self.iter_block(count)
                    iteration += count
                elif use_batch:
                    left = self.remaining()
                    count = gpu.config["batch"] * dist.world_rank()[0]
                    count = count if left is None else min(count, left)
                    # keep plot/save cadence of the reference: never step across a multiple of the step
                    for step in (plot_step, save_step):
                        if step:
                            count = min(count, step - iteration % step)
                    self.iter_batch(count)
                    iteration += count
                else:
                    self.iter()
                    iteration += 1
                self.process_output(iteration, plot_step=plot_step, save_step=save_step)
        except KeyboardInterrupt:
            pass
        finally:
            self.process_output(0, plot_step=plot_step, save_step=save_step)

    def process_output(self, iteration, plot_step, save_step):
        if plot_step and iteration % plot_step == 0:
            for result in self.results_list:
                result.plot_output()
            try:
                from IPython import display
                display.clear_output(wait=True)
            except ModuleNotFoundError:
                pass
        if save_step and iteration % save_step == 0 and dist.world_rank()[1] == 0:
            for result in self.results_list:
                result.save_output()
