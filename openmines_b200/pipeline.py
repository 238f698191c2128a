"""Streaming episode batches through one ``BatchedMine`` with the host copies off the critical path.

A parameter sweep (seeds x dispatchers, fleet ablations ...) runs many episode batches back to back and reads every
batch's production curves on the host.  Done synchronously, each batch pays its device->host copy (127 MB of per-minute
KPI records for 16384 north_pit instances) after the kernel.  ``EpisodePipeline`` double-buffers the outputs: batch k's
curves travel to pinned host memory on a copy stream while batch k+1 is being simulated, and ``collect()`` hands them
out in submission order.

    pipe = EpisodePipeline(mine, rng_mode="philox")
    for seeds in seed_blocks:
        pipe.submit(dispatcher_ids, seeds)          # host arrays; returns at once
        if pipe.pending > 1:
            summary, ticks = pipe.collect()         # numpy views of the oldest finished batch (valid until the next collect)
    while pipe.pending:
        summary, ticks = pipe.collect()
"""
import numpy as np

from .engine import RNG_MODES
from .trace import SUMMARY_DTYPE, TICK_DTYPE


class EpisodePipeline:
    def __init__(self, mine, rng_mode="philox", depth=2):
        torch = mine.torch
        self.mine, self.torch, self.mode = mine, torch, RNG_MODES[rng_mode]
        n, dev = mine.n, mine.device
        self.depth = int(depth)
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.h_seeds = [torch.empty(n, dtype=torch.int64).pin_memory() for _ in range(self.depth)]
        self.h_disp = [torch.empty(n, dtype=torch.int32).pin_memory() for _ in range(self.depth)]
        self.d_seeds = [torch.empty(n, dtype=torch.int64, device=dev) for _ in range(self.depth)]
        self.d_disp = [torch.empty(n, dtype=torch.int32, device=dev) for _ in range(self.depth)]
        self.results, self.h_summary, self.h_ticks = [], [], []
        self.computed = [torch.cuda.Event() for _ in range(self.depth)]     # kernel of slot s finished
        self.copied = [torch.cuda.Event() for _ in range(self.depth)]       # outputs of slot s are on the host
        self._slot_used = [False] * self.depth
        self._queue = []          # slots in submission order
        self._next = 0
        self.h2d_bytes = n * 12
        self.d2h_bytes = 0

    @property
    def pending(self):
        return len(self._queue)

    def submit(self, dispatcher_ids, seeds):
        """Enqueue one batch: upload (dispatcher, seed) per instance, reset, run, start the copy of its outputs."""
        torch, mine = self.torch, self.mine
        s = self._next
        if len(self._queue) == self.depth:
            raise RuntimeError("EpisodePipeline: collect() the oldest batch before submitting another (depth %d)" % self.depth)
        self._next = (s + 1) % self.depth
        main = torch.cuda.current_stream(mine.device)
        self.h_seeds[s].numpy()[:] = np.asarray(seeds).astype(np.uint64, copy=False).view(np.int64)
        self.h_disp[s].numpy()[:] = np.asarray(dispatcher_ids, dtype=np.int32)
        if self._slot_used[s]:
            main.wait_event(self.copied[s])                  # the previous outputs of this slot must have left the device
        self.d_seeds[s].copy_(self.h_seeds[s], non_blocking=True)
        self.d_disp[s].copy_(self.h_disp[s], non_blocking=True)
        mine.dispatcher_ids, mine.seeds = self.d_disp[s], self.d_seeds[s]
        mine.has_external = bool((self.h_disp[s].numpy() == 255).any())
        mine._ck(mine.lib.om_reset(mine.h, self.d_seeds[s].data_ptr(), self.d_disp[s].data_ptr(), self.mode,
                                   None, None, None, mine.state.data_ptr(), mine._stream()))
        if len(self.results) <= s:
            res = mine.run()
            self.results.append(res)
            self.h_summary.append(torch.empty(res.summary.shape, dtype=torch.uint8).pin_memory())
            self.h_ticks.append(torch.empty(res.ticks.shape, dtype=torch.uint8).pin_memory())
            self.d2h_bytes = self.h_summary[s].numel() + self.h_ticks[s].numel()
        else:
            mine.run(result=self.results[s])
        self.computed[s].record(main)
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.computed[s])
            self.h_summary[s].copy_(self.results[s].summary, non_blocking=True)
            self.h_ticks[s].copy_(self.results[s].ticks, non_blocking=True)
            self.copied[s].record(self.copy_stream)
        self._slot_used[s] = True
        self._queue.append(s)

    def collect(self):
        """Block until the oldest submitted batch is on the host; returns (OmSummary[n], OmTickRecord[n, T]) numpy views."""
        s = self._queue.pop(0)
        self.copied[s].synchronize()
        summary = self.h_summary[s].numpy().view(SUMMARY_DTYPE).reshape(-1)
        ticks = self.h_ticks[s].numpy().view(TICK_DTYPE).reshape(self.h_ticks[s].shape[0], -1)
        return summary, ticks
