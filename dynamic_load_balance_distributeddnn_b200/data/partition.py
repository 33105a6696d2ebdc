"""Dataset partitioning for unequal per-rank batch sizes.

Reference behaviour (``dataloader.py:12-49``): one global permutation from
``random.Random(seed)``, consecutive slices of ``int(frac·N)`` samples per rank, per-rank
batch ``B·frac``.  Two defects are fixed here: the reference derives each
rank's iteration count independently (``ceil(int(f·N)/int(B·f))``), which occasionally
differs between ranks and dead-locks the collectives; and it mixes float and int batch
sizes.  Here a single global ``steps_per_epoch = N // B`` is derived first and rank *r*
receives exactly ``steps · b_r`` samples, so every rank runs the same number of steps, each
global step consumes exactly ``B`` distinct samples, and the shards are disjoint.
"""
from __future__ import annotations

import random
from dataclasses import dataclass
from typing import List, Sequence

import numpy as np


def global_permutation(n: int, seed: int = 1234, shuffle: bool = True) -> np.ndarray:
    idx = list(range(n))
    if shuffle:
        rng = random.Random()
        rng.seed(seed)                      # same generator as the reference (dataloader.py:38-40)
        rng.shuffle(idx)
    return np.asarray(idx, dtype=np.int64)


@dataclass
class Shard:
    indices: np.ndarray        # sample ids owned by this rank this epoch
    local_batch: int
    steps: int

    def __len__(self) -> int:
        return len(self.indices)

    def batch_indices(self, step: int, order: np.ndarray) -> np.ndarray:
        sel = order[step * self.local_batch:(step + 1) * self.local_batch]
        return self.indices[sel]


class DataPartitioner:
    """Splits ``n`` samples across ranks for one epoch given integer local batches."""

    def __init__(self, n: int, local_batches: Sequence[int], seed: int = 1234, shuffle: bool = True,
                 max_steps: int = 0, start: int = 0):
        """``start``: number of samples of this epoch's permutation already consumed (mid-epoch re-partitioning with
        ``--rebalance_every``: the segments of one epoch still form a disjoint cover of the same permutation)."""
        self.n = int(n)
        self.local_batches = [int(b) for b in local_batches]
        self.global_batch = int(sum(self.local_batches))
        if self.global_batch <= 0:
            raise ValueError("global batch must be positive")
        self.start = int(start)
        self.steps = (self.n - self.start) // self.global_batch
        if max_steps > 0:
            self.steps = min(self.steps, max_steps)
        if self.steps == 0:
            raise ValueError(f"dataset of {n} samples is smaller than the global batch {self.global_batch}")
        perm = global_permutation(self.n, seed, shuffle)
        self.shards: List[Shard] = []
        off = self.start
        for b in self.local_batches:
            cnt = self.steps * b
            self.shards.append(Shard(perm[off:off + cnt], b, self.steps))
            off += cnt
        self.used = off

    def use(self, rank: int) -> Shard:
        return self.shards[rank]


def split_token_stream(n_tokens: int, local_batches: Sequence[int]) -> List[slice]:
    """LM variant: the token stream is cut *without shuffling* (reference ``dataloader.py:106``)
    into consecutive pieces ∝ local batch; each piece is then batchified into ``b_r`` columns.
    Every rank gets the same number of rows ``(n_tokens // B)`` so step counts agree."""
    total = int(sum(local_batches))
    rows = n_tokens // total
    out, off = [], 0
    for b in local_batches:
        out.append(slice(off, off + rows * int(b)))
        off += rows * int(b)
    return out
