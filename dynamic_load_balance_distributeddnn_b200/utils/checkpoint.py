"""Checkpoint / resume (absent from the reference — SURVEY §5.4).  Rank 0 writes
{model master weights, momentum, lr, epoch, reallocator state, RNG}; every rank loads it, so a run can
continue with the same partition vector and timing history."""
from __future__ import annotations

import os
from typing import Optional

import torch


def checkpoint_path(cfg) -> str:
    """One file per experiment, independent of the epoch budget (so ``-e 5`` can be resumed with ``-e 10``)."""
    import re
    stem = re.sub(r"-ep\d+", "", cfg.experiment_id("ckpt"))
    return os.path.join(cfg.checkpoint_dir, stem + ".pt")


def save_checkpoint(cfg, epoch: int, flat_state, reallocator, extra: Optional[dict] = None) -> str:
    os.makedirs(cfg.checkpoint_dir, exist_ok=True)
    path = checkpoint_path(cfg)
    blob = {"epoch": epoch, "flat": flat_state.state_dict(), "reallocator": reallocator.state_dict(),
            "torch_rng": torch.get_rng_state(), "extra": extra or {}}
    tmp = path + ".tmp"
    torch.save(blob, tmp)
    os.replace(tmp, path)
    return path


def load_checkpoint(cfg, flat_state, reallocator):
    path = checkpoint_path(cfg)
    if not os.path.isfile(path):
        return None
    blob = torch.load(path, map_location="cpu", weights_only=False)
    flat_state.load_state_dict(blob["flat"])
    reallocator.load_state_dict(blob["reallocator"])
    return blob
