"""nvidia-smi clock / throttle sampling during a timed region (bench.py's ``clocks`` key; recipe in
/opt/skills/guides/B200_PROFILING.md)."""
from __future__ import annotations

import statistics
import subprocess
import threading
from typing import Dict, List

_Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
      "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")


class ClockSampler:
    def __init__(self, gpu_index: int = 0, period_ms: int = 100):
        self.gpu_index, self.period_ms = gpu_index, period_ms
        self.samples: List[List[str]] = []
        self._proc = None
        self._thread = None

    def start(self) -> None:
        try:
            self._proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={_Q}", "--format=csv,noheader,nounits",
                                           "-i", str(self.gpu_index), "-lms", str(self.period_ms)],
                                          stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self._proc = None
            return

        def pump():
            for line in self._proc.stdout:
                parts = [p.strip() for p in line.strip().split(",")]
                if len(parts) >= 7:
                    self.samples.append(parts)
        self._thread = threading.Thread(target=pump, daemon=True)
        self._thread.start()

    def stop(self) -> Dict:
        if self._proc is not None:
            self._proc.terminate()
            try:
                self._proc.wait(timeout=2)
            except Exception:
                self._proc.kill()
            if self._thread is not None:
                self._thread.join(timeout=2)
        sm, mx, reasons = [], [], set()
        for s in self.samples:
            try:
                sm.append(float(s[0])); mx.append(float(s[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        busy = sorted(sm)[len(sm) // 2:]                  # upper half ≈ samples under load
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm)}
