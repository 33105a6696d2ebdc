"""Per-rank iteration-time tracker.

The reference measures ``compute = epoch wall − Σ time blocked in req.wait()`` with
``time.time()`` and no device sync (``dbs.py:222-250``, ``:297-299``), which on a GPU
mis-attributes asynchronous backward time to *sync* (SURVEY D10).  This tracker keeps
the same two quantities — compute seconds (the DBS feedback signal) and sync/straggler-wait
seconds — but sources them from the device:

* CUDA: a pair of CUDA events brackets every step's compute region on the compute stream;
  the straggler wait is measured *inside* the allreduce kernel's entry barrier with
  ``%globaltimer`` and accumulated in a device counter (``parallel/symm.py``), so no host
  sync is needed per step; events are resolved lazily at epoch end.
* CPU/gloo: ``time.perf_counter`` around the compute region and around the collective.

Injected straggle (host sleep or device burner placed between backward and the allreduce,
exactly where the reference puts ``fault_tolerance_wait``, ``dbs.py:236``) counts as compute.
"""
from __future__ import annotations

import time
from typing import List, Optional

import torch


class TimeTracker:
    def __init__(self, device: torch.device):
        self.device = torch.device(device)
        self.cuda = self.device.type == "cuda"
        self.reset()

    def reset(self) -> None:
        self._compute_s = 0.0
        self._sync_s = 0.0
        self._pairs: List = []
        self._unsteady_pairs: List = []
        self.unsteady_compute_s = 0.0
        self._t0: Optional[float] = None
        self._wall0 = time.perf_counter()
        self.steps = 0
        self.unsteady = 0

    # ---- compute region -------------------------------------------------------------
    def start_compute(self) -> None:
        if self.cuda:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            self._e0 = e
        else:
            self._t0 = time.perf_counter()

    def stop_compute(self, steady: bool = True) -> None:
        """``steady=False`` marks a step whose duration is not representative of this rank's throughput (eager warm-up
        before a CUDA graph exists for a new local batch size): it is excluded from the feedback signal and the epoch
        total is extrapolated from the steady steps, otherwise a rank that has just been re-sized looks slow, gets
        shrunk again, is re-sized again ... and the split runs away."""
        if not steady:
            # kept aside: only used when an epoch has NO steady step at all (e.g. max_steps_per_epoch <= warm-up steps)
            self.unsteady += 1
            if self.cuda:
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                self._unsteady_pairs.append((self._e0, e))
            else:
                self.unsteady_compute_s += time.perf_counter() - self._t0
            return
        if self.cuda:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            self._pairs.append((self._e0, e))
        else:
            self._compute_s += time.perf_counter() - self._t0
        self.steps += 1

    def add_compute(self, seconds: float) -> None:
        self._compute_s += seconds

    def add_sync(self, seconds: float) -> None:
        self._sync_s += seconds

    # ---- epoch summary -------------------------------------------------------------
    def finish(self):
        """→ (compute_seconds, sync_seconds, wall_seconds) for the epoch."""
        if self.cuda:
            torch.cuda.synchronize(self.device)
            for a, b in self._pairs:
                self._compute_s += a.elapsed_time(b) * 1e-3
            self._pairs.clear()
            for a, b in self._unsteady_pairs:
                self.unsteady_compute_s += a.elapsed_time(b) * 1e-3
            self._unsteady_pairs.clear()
        wall = time.perf_counter() - self._wall0
        if self.unsteady and self.steps:
            self._compute_s *= (self.steps + self.unsteady) / self.steps
        return self._compute_s, self._sync_s, wall
