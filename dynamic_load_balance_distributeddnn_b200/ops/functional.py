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
    if x.is_cuda and bias is None and weight.shape[1] <= 4:
        from . import stem as _stem
        if _stem.supported(x, weight, stride, padding, groups):
            return _stem.conv2d(x, weight)                  # RGB stem: direct SIMT kernel (K = 27 is below any tensor-core tile)
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


class _SoftmaxCEFn(torch.autograd.Function):
    """mean softmax cross-entropy: loss and d(loss)/d(logits) from ONE kernel (``csrc/loss.cu``); the gradient seed is
    multiplied by ``grad_scale`` (a device scalar: the rank's DBS weight) inside that kernel."""

    @staticmethod
    def forward(ctx, logits, target, grad_scale):
        lib = nat.require()
        b, c = logits.shape
        lg = logits if logits.stride(1) == 1 else logits.contiguous()
        dlog = torch.empty((b, c), dtype=torch.float32, device=logits.device)
        loss = torch.empty((), dtype=torch.float32, device=logits.device)
        nat.check(lib.dlb_softmax_ce_small(nat.dtype_code(lg.dtype), lg.data_ptr(), lg.stride(0), target.data_ptr(), dlog.data_ptr(),
                                           loss.data_ptr(), nat.ptr(grad_scale), b, c, nat.stream_ptr(logits.device)), "softmax_ce")
        ctx.save_for_backward(dlog)
        ctx.in_dtype = logits.dtype
        return loss

    @staticmethod
    def backward(ctx, go):
        (dlog,) = ctx.saved_tensors
        return (dlog * go).to(ctx.in_dtype), None, None


def cross_entropy(logits: torch.Tensor, target: torch.Tensor, grad_scale: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Mean cross-entropy in fp32 regardless of the logits dtype.  ``grad_scale`` (device scalar) multiplies the gradient
    seed only -- the returned loss value is unscaled."""
    if (logits.is_cuda and nat.available() and logits.dim() == 2 and logits.shape[1] <= 1024 and target.dtype == torch.int64
            and logits.dtype in (torch.float32, torch.bfloat16) and hasattr(nat.get(), "dlb_softmax_ce_small")):
        return _SoftmaxCEFn.apply(logits, target.contiguous(), grad_scale)
    loss = F.cross_entropy(logits.float(), target)
    if grad_scale is not None and loss.requires_grad:
        # value unchanged, gradient scaled:  loss + (s - 1) * (loss - loss.detach())
        loss = loss.detach() + grad_scale.reshape(()).to(loss.dtype) * (loss - loss.detach())
    return loss


def nll_loss(log_probs: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    return F.nll_loss(log_probs.float(), target)
