"""Functional op surface used by the model zoo.

Every op has a plain-PyTorch reference path (CPU debug mode, numerics tests) and, where a
hand-written sm_100a kernel exists, dispatches to it on CUDA.  Convolutions take/return
logical-NCHW tensors in channels_last (NHWC) memory on CUDA so the NHWC kernels and the
TMA/tcgen05 implicit-GEMM path see their native layout.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn.functional as F

from . import _native as nat
from .norm import group_norm_act  # noqa: F401  (re-export)


def conv2d(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None, stride: int = 1,
           padding: int = 0, groups: int = 1) -> torch.Tensor:
    """2-D convolution.  bf16/fp32; dispatches to the tcgen05 implicit-GEMM kernel when the shape is
    supported (``ops/conv.py``), else to the vendor library through ATen."""
    from . import conv as _conv
    if x.is_cuda and _conv.tc_supported(x, weight, stride, padding, groups):
        return _conv.conv2d_tc(x, weight, bias, stride, padding)
    if bias is not None and bias.dtype != x.dtype:
        bias = bias.to(x.dtype)
    if weight.dtype != x.dtype:
        weight = weight.to(x.dtype)
    return F.conv2d(x, weight, bias, stride=stride, padding=padding, groups=groups)


def linear(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``F.linear``; bf16 shapes the tcgen05 GEMM supports (K, N multiples of 8, >= 128 rows) run on it."""
    if x.is_cuda and weight.dtype == x.dtype:
        from . import conv as _conv
        if _conv._ENABLED and nat.available():
            from . import gemm_tc
            if gemm_tc.linear_supported(x, weight):
                return gemm_tc.linear(x, weight, bias)
    if weight.dtype != x.dtype:
        weight = weight.to(x.dtype)
    if bias is not None and bias.dtype != x.dtype:
        bias = bias.to(x.dtype)
    return F.linear(x, weight, bias)


def cross_entropy(logits: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """Mean cross-entropy in fp32 regardless of the logits dtype."""
    return F.cross_entropy(logits.float(), target)


def nll_loss(log_probs: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    return F.nll_loss(log_probs.float(), target)
