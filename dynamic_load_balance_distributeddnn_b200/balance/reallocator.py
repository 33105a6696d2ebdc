"""The DBS reallocator: re-split a fixed global batch ``B`` across ranks in proportion to
each rank's measured throughput.

Algorithm (reference ``dbs.py:458-476``): with previous shares ``p_i``
and pure-compute times ``t_i`` of the last epoch, rank *i* processed ``p_i / t_i`` share per
second, so its new share is ``r_i = (p_i/t_i) / Σ_j (p_j/t_j)``.  The integer local batches
are ``round(r_i · B)`` under some rounding rule, and the shares handed to the data
partitioner / gradient weights are ``b_i / Σ b``.

Rounding rules
  * ``largest_remainder`` (default): floor, then hand the ``B − Σfloor`` leftover samples to
    the largest fractional parts.  Σ b_i == B always (fixes SURVEY D5), every rank keeps at
    least ``min_local`` samples so a slow rank is never absorbing (fixes D7), and batches can be
    constrained to a multiple of ``quantum`` (CUDA-graph / tile friendly).
  * ``reference``: the reference's rule — only fractional parts that are both among the
    ``B − Σfloor`` largest *and* ≥ 0.5 are rounded up (``dbs.py:469-473``), so Σ b_i may be
    ``B−1`` or ``B−2``.  Kept for bit-parity experiments.
"""
from __future__ import annotations

from typing import Sequence, Tuple

import numpy as np


def throughput_shares(nodes_time: Sequence[float], partition: Sequence[float]) -> np.ndarray:
    t = np.maximum(np.asarray(nodes_time, dtype=np.float64), 1e-12)
    p = np.asarray(partition, dtype=np.float64)
    rate = p / t
    s = rate.sum()
    if not np.isfinite(s) or s <= 0:
        return np.full(len(p), 1.0 / len(p))
    return rate / s


def _largest_remainder(target: np.ndarray, total: int) -> np.ndarray:
    """Integers ≥ 0 summing to ``total`` closest (in the Hamilton sense) to ``target``."""
    fl = np.floor(target + 1e-12).astype(np.int64)
    left = int(total - fl.sum())
    if left > 0:
        frac = np.round(target - fl, 9)           # remainders equal up to float noise are ties
        # ties go to the larger share, then to the lower rank index: every rank computes the same answer and the multiset
        # of batches does not depend on the order of the ranks
        order = np.lexsort((np.arange(len(frac)), -target, -frac))
        fl[order[:left]] += 1
    elif left < 0:                      # only reachable through min/quantum clamping
        frac = target - fl
        order = np.lexsort((np.arange(len(frac)), frac))
        k = 0
        while left < 0:
            i = order[k % len(order)]
            if fl[i] > 0:
                fl[i] -= 1
                left += 1
            k += 1
    return fl


def integer_split(shares: Sequence[float], batch_size: int, min_local: int = 1,
                  quantum: int = 1) -> np.ndarray:
    """Split ``batch_size`` into per-rank integers ∝ ``shares`` with Σ == batch_size,
    each ≥ ``min_local`` (when feasible) and a multiple of ``quantum`` (when feasible)."""
    shares = np.asarray(shares, dtype=np.float64)
    n = len(shares)
    shares = shares / shares.sum()
    q = max(1, int(quantum))
    if batch_size % q != 0 or batch_size // q < n:
        q = 1
    units = batch_size // q
    min_units = max(0, -(-int(min_local) // q))          # ceil(min_local / q)
    if min_units * n > units:
        min_units = units // n
    # water-filling: pin ranks that fall under the minimum, re-split the rest proportionally
    fixed = np.zeros(n, dtype=bool)
    out = np.zeros(n, dtype=np.float64)
    for _ in range(n + 1):
        free_units = units - min_units * int(fixed.sum())
        free_share = shares[~fixed].sum()
        if free_share <= 0:
            out[~fixed] = free_units / max(1, int((~fixed).sum()))
        else:
            out[~fixed] = shares[~fixed] / free_share * free_units
        out[fixed] = min_units
        newly = (~fixed) & (out < min_units - 1e-9)
        if not newly.any():
            break
        fixed |= newly
    ints = np.empty(n, dtype=np.int64)
    ints[fixed] = min_units
    if (~fixed).any():
        free_total = units - min_units * int(fixed.sum())
        sub = _largest_remainder(out[~fixed], free_total)
        # rounding may still push a free rank below the minimum by one unit; repair
        ints[~fixed] = sub
        low = (~fixed) & (ints < min_units)
        while low.any():
            i = int(np.argmax(low))
            j = int(np.argmax(np.where(fixed | low, -1, ints)))
            ints[i] += 1
            ints[j] -= 1
            low = (~fixed) & (ints < min_units)
    return ints * q


def reference_split(shares: Sequence[float], batch_size: int) -> np.ndarray:
    """Bit-compatible with the reference's rounding (``dbs.py:465-473``), including its
    sample loss and the ``argsort()[-0:]`` quirk (D6)."""
    shares = np.asarray(shares, dtype=np.float64)
    norm_batch = shares * batch_size / shares.sum()
    fl = np.floor(norm_batch)
    ceil_counter = int(batch_size - int(fl.sum()))
    frac = norm_batch - fl
    idx_ceil = frac.argsort()[-ceil_counter:]           # ceil_counter == 0 ⇒ all indices (D6)
    idx_round = np.argwhere(frac >= 0.5).reshape(-1)
    idx = np.intersect1d(idx_ceil, idx_round)
    fl[idx] += 1
    return fl.astype(np.int64)


def get_size(nodes_time: Sequence[float], partition_size: Sequence[float], batch_size: int,
             rounding: str = "largest_remainder", min_local: int = 1, quantum: int = 1
             ) -> Tuple[np.ndarray, np.ndarray]:
    """New ``(fractions, local_batches)`` from last epoch's compute times and shares.

    ``fractions = local_batches / Σ local_batches`` — exactly what the reference returns
    (``dbs.py:474``) and what feeds the data partitioner and the allreduce weights."""
    r = throughput_shares(nodes_time, partition_size)
    if rounding == "reference":
        ints = reference_split(r, batch_size)
    else:
        ints = integer_split(r, batch_size, min_local=min_local, quantum=quantum)
    tot = ints.sum()
    frac = ints / tot if tot > 0 else np.full(len(ints), 1.0 / len(ints))
    return frac, ints


class Reallocator:
    """Stateful wrapper used by the trainer: owns the current split, applies an optional EMA
    to the time signal, and only moves when DBS is enabled (``-dbs``; static DP otherwise,
    reference ``dbs.py:379,388``)."""

    def __init__(self, world_size: int, batch_size: int, enabled: bool = True,
                 rounding: str = "largest_remainder", min_local: int = 1, quantum: int = 1,
                 ema: float = 0.0):
        self.world_size = world_size
        self.batch_size = batch_size
        self.enabled = enabled
        self.rounding = rounding
        self.min_local = min_local
        self.quantum = quantum
        self.ema = ema
        self.nodes_time = np.ones(world_size, dtype=np.float64)          # dbs.py:378
        self.fractions = np.full(world_size, 1.0 / world_size)           # dbs.py:379
        self.local_batches = self._initial_split()
        self.history = []
        self._have_obs = False            # nodes_time still holds the all-ones placeholder: nothing to blend an EMA with

    def _initial_split(self) -> np.ndarray:
        if self.rounding == "reference":
            return np.full(self.world_size, int(self.batch_size * (1.0 / self.world_size)), dtype=np.int64)
        return integer_split(np.ones(self.world_size), self.batch_size, self.min_local, self.quantum)

    def observe(self, nodes_time: Sequence[float]) -> None:
        t = np.asarray(nodes_time, dtype=np.float64)
        if self.ema > 0 and self._have_obs:
            t = self.ema * self.nodes_time + (1.0 - self.ema) * t
        self.nodes_time = t
        self._have_obs = True

    def step(self) -> Tuple[np.ndarray, np.ndarray]:
        """Called at the start of every epoch (or every N steps).  Returns the split to use."""
        if self.enabled:
            self.fractions, self.local_batches = get_size(
                self.nodes_time, self.fractions, self.batch_size, self.rounding,
                self.min_local, self.quantum)
        self.history.append(self.local_batches.copy())
        return self.fractions, self.local_batches

    def weights(self, uniform: bool = False) -> np.ndarray:
        """Gradient-allreduce weights ``w_r = b_r / Σ b`` (reference ``dbs.py:293``), or the
        ``-de`` ablation ``1/n``."""
        if uniform:
            return np.full(self.world_size, 1.0 / self.world_size)
        return self.fractions / self.fractions.sum()

    def state_dict(self):
        return {"nodes_time": self.nodes_time.tolist(), "fractions": self.fractions.tolist(),
                "local_batches": self.local_batches.tolist(), "have_obs": bool(self._have_obs)}

    def load_state_dict(self, sd):
        self.nodes_time = np.asarray(sd["nodes_time"], dtype=np.float64)
        self.fractions = np.asarray(sd["fractions"], dtype=np.float64)
        self.local_batches = np.asarray(sd["local_batches"], dtype=np.int64)
        self._have_obs = bool(sd.get("have_obs", True))


class AffineReallocator(Reallocator):
    """DBS with a per-rank *affine* step-time model  t_r(b) = alpha_r + beta_r * b.

    The reference's rule (``get_size``) assumes time is proportional to the local batch.  On a B200 with small per-rank
    batches a step is launch-latency bound — its time barely depends on the batch (measured: 9.7 ms at b = 64 and b = 128)
    — and a straggler's extra time is largely a fixed cost: the proportional rule then keeps shrinking the slow rank
    geometrically down to the minimum batch although that buys nothing.  This variant fits (alpha_r, beta_r) from the
    history of (local batch, compute time) observations (last ``window`` distinct batch sizes per rank; a slope pooled over
    all ranks with per-rank intercepts, each rank's own slope shrunk towards it; proportional rule for the very first move,
    when nothing is identifiable yet) and picks the split that equalises the predicted step
    times:  b_r = (tau - alpha_r) / beta_r  with  tau  such that  sum_r b_r = B  (water-filling at the lower bound).
    A rank whose time does not respond to its batch at all (flat fit) keeps its batch.
    ``--dbs_model affine``; ``auto`` (the default) selects it on CUDA devices and the reference's rule on CPU.
    """

    def __init__(self, *args, window: int = 6, noise_frac: float = 0.05, **kw):
        super().__init__(*args, **kw)
        self.window = window
        self.noise_frac = noise_frac
        self.obs = [[] for _ in range(self.world_size)]          # per rank: list of (batch, time)

    def observe(self, nodes_time: Sequence[float]) -> None:
        super().observe(nodes_time)
        for r, t in enumerate(self.nodes_time):
            b = float(self.local_batches[r])
            same = [p for p in self.obs[r] if p[0] == b]
            t = float(t) if not same else 0.5 * (same[-1][1] + float(t))   # one (smoothed) point per distinct batch size
            self.obs[r] = [p for p in self.obs[r] if p[0] != b][-(self.window - 1):] + [(b, t)]

    def state_dict(self):
        sd = super().state_dict()
        sd["obs"] = [[list(map(float, p)) for p in per_rank] for per_rank in self.obs]      # the (batch, time) history the fit needs
        return sd

    def load_state_dict(self, sd):
        super().load_state_dict(sd)
        if "obs" in sd:
            self.obs = [[(float(b), float(t)) for b, t in per_rank] for per_rank in sd["obs"]]

    def _pooled_slope(self):
        """Within-rank least-squares slope over ALL ranks' observations (fixed-effects model: one intercept per rank, one common
        slope).  The ranks of a job are the same kind of GPU running the same kernels, so the marginal cost of a sample is (nearly)
        common while the fixed cost differs (a straggler's injected / contended milliseconds).  Pooling is what makes the model
        identifiable early: after the first proportional move the fast ranks have only moved by a sample or two -- their own
        two-point slopes are noise -- but the straggler's large move pins the common slope."""
        sxx = sxy = 0.0
        for pts in self.obs:
            if len(pts) < 2:
                continue
            bs = np.array([q[0] for q in pts]); ts = np.array([q[1] for q in pts])
            sxx += float(((bs - bs.mean()) ** 2).sum())
            sxy += float(((bs - bs.mean()) * (ts - ts.mean())).sum())
        if sxx < 1.0:
            return None
        # Ridge towards the proportional model's slope (time / batch, i.e. no fixed cost): while the batches have only moved by a
        # few samples the data cannot tell a 0.04 ms/sample slope from timing noise (measured at 2 GPUs: +-0.4 ms on 17 ms with
        # batches 256 / 268 / 273 gave a slope 5x too small and a 160-sample jump), so the estimate starts at the reference's
        # rule -- same step size to first order -- and hands over to the data as the spread grows past a quarter of the mean
        # local batch.
        bs_all = np.array([q[0] for pts in self.obs for q in pts]); ts_all = np.array([q[1] for pts in self.obs for q in pts])
        prop = float(ts_all.mean() / max(bs_all.mean(), 1e-9))
        kappa = (0.25 * self.batch_size / self.world_size) ** 2
        return (sxy + kappa * prop) / (sxx + kappa)

    def _fit(self, r: int, pooled):
        """(alpha_r, beta_r): the rank's own slope shrunk towards the pooled one (prior weight = one pair of observations a
        quarter of the mean batch apart), intercept from the rank's mean point."""
        pts = self.obs[r]
        if not pts or pooled is None:
            return None
        bs = np.array([q[0] for q in pts]); ts = np.array([q[1] for q in pts])
        sxx = float(((bs - bs.mean()) ** 2).sum())
        sxy = float(((bs - bs.mean()) * (ts - ts.mean())).sum())
        kappa = (0.25 * self.batch_size / self.world_size) ** 2
        beta = (sxy + kappa * pooled) / (sxx + kappa)
        if not np.isfinite(beta) or beta * bs.mean() < 0.02 * ts.mean():
            return float(ts.mean()), 0.0                          # flat: the batch does not move this rank's time
        alpha = float(ts.mean() - beta * bs.mean())
        if alpha < 0.0:
            # a negative fixed cost is not physical: the borrowed slope is too steep for this rank (e.g. the pooled slope is
            # dominated by a genuinely slower device) -> the reference's proportional model for this rank
            return 0.0, float(ts.mean() / bs.mean())
        return alpha, float(beta)

    def _typical_ranks(self, pooled):
        """The ranks of one job are the same devices running the same kernels: fixed-cost differences of a few percent of a
        step between them are measurement noise (5 warm steps per observation; collectives waiting for a late peer share the SMs
        with the backward pass), and chasing them costs more than it can gain -- a 0.5 ms error moves 10 samples at
        0.047 ms/sample.  Ranks whose fixed cost under the pooled slope lies within `noise_frac` of a step of the median form ONE
        group with one line (mean intercept, pooled slope) and therefore get equal batches; the others -- real stragglers are
        tens of percent off -- keep their own fit.  Returns (mask, group intercept) or (None, None)."""
        if self.world_size < 3 or self.noise_frac <= 0 or pooled is None or any(not p for p in self.obs):
            return None, None
        bm = np.array([np.mean([q[0] for q in pts]) for pts in self.obs])
        tm = np.array([np.mean([q[1] for q in pts]) for pts in self.obs])
        a = tm - pooled * bm
        tau = self.noise_frac * float(np.median(tm))
        mask = np.abs(a - np.median(a)) <= tau
        if mask.sum() < 2:
            return None, None
        return mask, float(a[mask].mean())

    def step(self) -> Tuple[np.ndarray, np.ndarray]:
        if not self.enabled:
            return super().step()
        pooled = self._pooled_slope()
        fits = [self._fit(r, pooled) for r in range(self.world_size)]
        if any(f is None for f in fits):
            return super().step()                                 # not identifiable yet: proportional (reference) rule
        typical, a_typ = self._typical_ranks(pooled)
        if typical is not None and a_typ >= 0.0 and pooled > 0.0:
            fits = [(a_typ, float(pooled)) if typical[r] else fits[r] for r in range(self.world_size)]
        alpha = np.array([f[0] for f in fits]); beta = np.array([f[1] for f in fits])
        lo = float(max(1, self.min_local))
        b = self.local_batches.astype(np.float64).copy()
        active = beta > 0.0                                       # flat ranks keep their batch: moving it buys nothing
        if active.sum() >= 2:
            pinned = np.zeros(self.world_size, dtype=bool)        # ranks water-filled at the lower bound
            for _ in range(self.world_size + 1):
                live = active & ~pinned
                if not live.any():
                    break
                budget = self.batch_size - b[~active].sum() - lo * pinned.sum()
                tau = (budget + (alpha[live] / beta[live]).sum()) / (1.0 / beta[live]).sum()
                b[live] = (tau - alpha[live]) / beta[live]
                b[pinned] = lo
                under = live & (b < lo)
                if not under.any():
                    break
                pinned |= under
        self.local_batches = integer_split(np.maximum(b, 1e-9), self.batch_size, self.min_local, self.quantum)
        self.fractions = self.local_batches / self.local_batches.sum()
        self.history.append(self.local_batches.copy())
        return self.fractions, self.local_batches
