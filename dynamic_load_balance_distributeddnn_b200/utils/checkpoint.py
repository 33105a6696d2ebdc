"""Checkpoint / resume (absent from the reference — SURVEY §5.4).  Rank 0 writes
{model master weights, momentum, lr, epoch, reallocator state (incl. the affine model's (batch, time) history), torch RNG
state, step counters, stats history}; every rank loads it, so a run continues with the same partition vector, timing
history, augmentation stream and statistics lists.  The flat buffers are stored WITHOUT the world-size-dependent tail padding,
so a checkpoint written at one world size resumes at another."""
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
    rsd = blob["reallocator"]
    if len(rsd.get("nodes_time", [])) == reallocator.world_size:      # a different world size starts from the uniform split
        reallocator.load_state_dict(rsd)
    if blob.get("torch_rng") is not None:
        torch.set_rng_state(blob["torch_rng"])
    return blob
