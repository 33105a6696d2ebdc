"""Learning-rate policies behind ``-ocp``.

``legacy`` reproduces the curve that is actually live in the reference (``dbs.py:193-215``): the
warm-up leg is commented out and the decay term uses ``0.7*epoch`` instead of ``0.7*epoch_size``, so
for ``epoch ≥ 0.7·E`` the rate is ``lr·(1 − 0.99·epoch/E)`` — a discontinuous drop.
``one_cycle`` (default) is the policy its docstring describes: linear warm-up 0.01·lr→lr over the first
30 %, plateau, linear decay lr→0.01·lr over the last 30 %.  Both are disabled by ``-de``
(``dbs.py:202-203``).
"""
from __future__ import annotations


def lr_at_epoch(base_lr: float, epoch: int, epoch_size: int, policy: str = "one_cycle", enabled: bool = True,
                disabled_enhancements: bool = False) -> float:
    if not enabled or disabled_enhancements or epoch_size <= 0:
        return base_lr
    e, n = float(epoch), float(epoch_size)
    if policy == "legacy":
        if 0.7 * n <= e < n:
            return base_lr - ((0.99 * base_lr) / (0.3 * n)) * (e - 0.7 * e)
        return base_lr
    lo = 0.01 * base_lr
    if e < 0.3 * n:
        return lo + (base_lr - lo) * (e / (0.3 * n))
    if e < 0.7 * n:
        return base_lr
    return base_lr - (base_lr - lo) * min(1.0, (e - 0.7 * n) / (0.3 * n))
