"""Native (sm_100a) fast paths for the language-model ops; each ``has_*`` reports whether the
kernel is present in the built library so ``transformer_ops`` can fall back explicitly."""
from __future__ import annotations

from . import _native as nat


def has_add_layer_norm() -> bool:
    lib = nat.get()
    return lib is not None and hasattr(lib, "dlb_add_layer_norm_fwd")


def has_linear_ce() -> bool:
    return False


def add_layer_norm(x, residual, weight, bias, eps):
    from .layernorm import add_layer_norm as f
    return f(x, residual, weight, bias, eps)
