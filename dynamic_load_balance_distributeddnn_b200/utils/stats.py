"""Rank-0 statistics recorder: the reference's ``./statis/<id>.npy`` pickled dict of per-epoch lists
(``dbs.py:316-326,428-442``) plus a JSON twin and extra device-side metrics.  The
directory is created (the reference never creates it and fails at the final ``np.save``, D2)."""
from __future__ import annotations

import json
import os
from typing import Dict, List

import numpy as np

REFERENCE_KEYS = ("epoch", "train_loss", "train_time", "sync_time", "val_loss", "accuracy", "partition",
                  "node_time", "wallclock_time")
EXTRA_KEYS = ("local_batches", "samples_per_sec", "straggler_wait_ms_per_step", "steps", "lr")


class StatsRecorder:
    def __init__(self, cfg, enabled: bool = True):
        self.cfg = cfg
        self.enabled = enabled
        self.data: Dict[str, List] = {k: [] for k in REFERENCE_KEYS + EXTRA_KEYS}

    def append(self, **kw) -> None:
        if not self.enabled:
            return
        for k in self.data:
            if k in kw:
                self.data[k].append(kw[k])

    def state(self) -> dict:
        """History so far (saved in checkpoints so a resumed run continues the same per-epoch lists)."""
        return {k: list(v) for k, v in self.data.items()}

    def load_state(self, sd: dict) -> None:
        for k in self.data:
            if k in sd:
                self.data[k] = list(sd[k])

    def path(self, ext: str = ".npy") -> str:
        return os.path.join(self.cfg.stats_dir, self.cfg.experiment_id(0) + ext)

    def save(self) -> str:
        if not self.enabled:
            return ""
        os.makedirs(self.cfg.stats_dir, exist_ok=True)
        np.save(self.path(".npy"), self.data, allow_pickle=True)

        def js(v):
            if isinstance(v, np.ndarray):
                return v.tolist()
            if isinstance(v, (np.floating, np.integer)):
                return v.item()
            if isinstance(v, (list, tuple)):
                return [js(x) for x in v]
            return v
        with open(self.path(".json"), "w") as f:
            json.dump({k: js(v) for k, v in self.data.items()}, f)
        return self.path(".npy")


def load_stats(path: str) -> dict:
    return np.load(path, allow_pickle=True).item()
