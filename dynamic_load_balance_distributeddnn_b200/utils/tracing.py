"""Tracing / profiling (``--profile true``).

The reference has no profiling hooks at all (wall-clock bookkeeping only, ``dbs.py:226,297-299``).
With ``--profile`` every rank gets

* **NVTX ranges** around the phases of a step (``stage_h2d``, ``augment``, ``forward``, ``backward``, ``straggler``,
  ``reduce_and_step``) and of an epoch (``rebalance``, ``validate``): they show up as ranges in ``ncu``/``nsys``
  and let ``ncu --nvtx --nvtx-include "forward/"`` capture one phase;
* a **torch.profiler capture** of a few steady-state steps → ``<log_dir>/<experiment id>.trace.json`` (Chrome /
  Perfetto trace) and ``<experiment id>.kernels.txt`` (per-kernel totals, the table the launch summaries under
  ``profiles/`` are the ncu counterpart of);
* a **host-side phase table** (wall-clock seconds the issuing thread spent per phase) in the rank log at the end.

Profiling runs execute the step eagerly (no CUDA graph: a replayed graph is one opaque launch) — they are for
attribution, never for throughput numbers.  The device-side ``%globaltimer`` stamps that feed the balancer
(``Trainer._stamp_*``) are always on and independent of this module.
"""
from __future__ import annotations

import contextlib
import os
import time
from collections import OrderedDict
from typing import Optional

import torch


class Tracer:
    def __init__(self, enabled: bool, stem: Optional[str] = None, cuda: bool = False, skip_steps: int = 4,
                 active_steps: int = 4, logger=None):
        self.enabled = bool(enabled)
        self.stem = stem
        self.cuda = bool(cuda) and torch.cuda.is_available()
        self.logger = logger
        self.phase_s: "OrderedDict[str, float]" = OrderedDict()
        self.phase_n: "OrderedDict[str, int]" = OrderedDict()
        self._prof = None
        self.trace_path = self.table_path = None
        if self.enabled and stem:
            self.trace_path, self.table_path = stem + ".trace.json", stem + ".kernels.txt"
            acts = [torch.profiler.ProfilerActivity.CPU]
            if self.cuda:
                acts.append(torch.profiler.ProfilerActivity.CUDA)
            self._prof = torch.profiler.profile(
                activities=acts, schedule=torch.profiler.schedule(wait=0, warmup=skip_steps, active=active_steps, repeat=1),
                on_trace_ready=self._export, record_shapes=False, with_stack=False)
            self._prof.__enter__()

    # ---- phases ----------------------------------------------------------------------------------------
    @contextlib.contextmanager
    def range(self, name: str):
        if not self.enabled:
            yield
            return
        if self.cuda:
            torch.cuda.nvtx.range_push(name)
        t0 = time.perf_counter()
        try:
            with torch.profiler.record_function(name):
                yield
        finally:
            dt = time.perf_counter() - t0
            if self.cuda:
                torch.cuda.nvtx.range_pop()
            self.phase_s[name] = self.phase_s.get(name, 0.0) + dt
            self.phase_n[name] = self.phase_n.get(name, 0) + 1

    def step(self) -> None:
        """Call once per optimisation step: advances the torch.profiler schedule."""
        if self._prof is not None:
            self._prof.step()

    # ---- output ----------------------------------------------------------------------------------------
    def _export(self, prof) -> None:
        os.makedirs(os.path.dirname(os.path.abspath(self.trace_path)), exist_ok=True)
        prof.export_chrome_trace(self.trace_path)
        sort_key = "self_cuda_time_total" if self.cuda else "self_cpu_time_total"
        try:
            table = prof.key_averages().table(sort_by=sort_key, row_limit=60)
        except Exception:                                   # older/newer kineto builds name the column differently
            table = prof.key_averages().table(row_limit=60)
        with open(self.table_path, "w") as f:
            f.write(table)

    def phase_table(self) -> str:
        rows = [f"{'phase':<18}{'calls':>8}{'host s':>12}{'ms/call':>12}"]
        for k, s in self.phase_s.items():
            n = self.phase_n[k]
            rows.append(f"{k:<18}{n:>8}{s:>12.4f}{1e3 * s / max(1, n):>12.3f}")
        return "\n".join(rows)

    def close(self) -> None:
        if self._prof is not None:
            try:
                self._prof.__exit__(None, None, None)
            finally:
                self._prof = None
        if self.enabled and self.logger is not None and self.phase_s:
            self.logger.info("host-side phase table (issuing thread, wall clock):\n" + self.phase_table())
