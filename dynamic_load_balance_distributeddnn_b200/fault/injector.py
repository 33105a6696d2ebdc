"""Straggler ("fault-tolerance") injection.

The reference's injector (``dbs.py:94-129``): once per epoch each worker draws ``luck``; with
probability ``ftc`` it becomes slow for ``U{4..20}`` epochs, sleeping ``U{5..10}s / num_batches``
after every backward — *between backward and the allreduce* (``dbs.py:236,273``) so the delay counts
as that rank's compute time and DBS shrinks its batch.  The reference version crashes with a
``NameError`` as soon as ``-ft true`` is used (``saved_epoch`` is never defined); this is
the intended behaviour, plus deterministic modes the GPU experiments need (one rank per GPU makes the
reference's other trick — oversubscribing a GPU with ``-gpu 0,0,0,1`` — unavailable):

* ``throttle_rank`` / ``throttle_ms``: rank r is slowed by a fixed x ms per step,
* ``mode='burn'``: the delay is a device-side spin kernel on the compute stream (CUDA-graph
  capturable, shows up in device timers) instead of a host sleep.
"""
from __future__ import annotations

import random
import time
from typing import Optional

import torch

from ..ops import _native as nat


class StragglerInjector:
    def __init__(self, rank: int, enabled: bool = False, chance: float = 0.1, throttle_rank: int = -1,
                 throttle_ms: float = 0.0, mode: str = "sleep", device: Optional[torch.device] = None,
                 seed: Optional[int] = None, logger=None):
        self.rank, self.enabled, self.chance = rank, enabled, chance
        self.mode = mode
        self.device = torch.device(device) if device is not None else torch.device("cpu")
        self.logger = logger
        self.rng = random.Random(seed) if seed is not None else random.Random()   # reference: unseeded ⇒ per-rank
        self.fixed_ms = throttle_ms if (throttle_rank == rank and throttle_ms > 0) else 0.0
        # random-phase state (reference globals fault_wait / fault_round / fault_wait_time)
        self.waiting = False
        self.until_epoch = 0
        self.wait_seconds = 0.0
        self._seen_epoch = -1
        self._per_step_s = 0.0
        self.usec_t = None
        if self.device.type == "cuda":
            self.usec_t = torch.zeros(1, dtype=torch.float32, device=self.device)

    # ---- per epoch ------------------------------------------------------------------------------
    def begin_epoch(self, epoch: int, num_batches: int) -> float:
        """Decide this epoch's per-step delay (seconds).  Called once per epoch on every rank."""
        per_step = self.fixed_ms * 1e-3
        if self.enabled and epoch != self._seen_epoch:
            self._seen_epoch = epoch
            if self.waiting and epoch > self.until_epoch:
                self.waiting = False
            if not self.waiting:
                luck = self.rng.random()
                if self.logger:
                    self.logger.info(f"Rank {self.rank} got a luck of {luck}, limit is {self.chance}")
                if luck < self.chance:
                    self.wait_seconds = float(self.rng.randint(5, 10))
                    self.until_epoch = epoch + self.rng.randint(4, 20)
                    self.waiting = True
                    if self.logger:
                        self.logger.info(f"Rank {self.rank} starts to have a {self.wait_seconds} seconds more "
                                         f"waiting until epoch {self.until_epoch} !")
        if self.enabled and self.waiting:
            per_step += self.wait_seconds / float(max(1, num_batches))
        self._per_step_s = per_step
        if self.usec_t is not None:
            self.usec_t.fill_(per_step * 1e6 if self.mode == "burn" else 0.0)
        return per_step

    @property
    def per_step_seconds(self) -> float:
        return self._per_step_s

    @property
    def active(self) -> bool:
        return self._per_step_s > 0

    # ---- per step -------------------------------------------------------------------------------
    def host_delay(self) -> float:
        """Host-sleep flavour; returns the seconds slept (to be booked as compute time)."""
        if self.mode == "sleep" and self._per_step_s > 0:
            time.sleep(self._per_step_s)
            return self._per_step_s
        return 0.0

    def device_delay(self) -> None:
        """Device-burner flavour: enqueue the spin kernel on the current stream (graph-capturable; the
        duration is read from device memory so it can change between replays)."""
        if self.mode == "burn" and self.usec_t is not None and nat.available():
            nat.check(nat.require().dlb_burn(self.usec_t.data_ptr(), 0.0, nat.stream_ptr(self.device)), "burn")
