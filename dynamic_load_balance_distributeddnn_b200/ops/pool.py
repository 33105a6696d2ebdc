"""Pooling on channels-last tensors (``csrc/pool.cu``): non-overlapping average pooling (incl. the global pool of the
classifier heads) and general max pooling, with PyTorch fallbacks."""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import _native as nat


class _AvgPoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, k):
        n, c, h, w = x.shape
        x = x.contiguous(memory_format=torch.channels_last)
        y = torch.empty((n, c, h // k, w // k), dtype=x.dtype, device=x.device, memory_format=torch.channels_last)
        nat.check(nat.require().dlb_avgpool_nhwc(0, nat.dtype_code(x.dtype), x.data_ptr(), y.data_ptr(), n, h, w, c, k,
                                                 nat.stream_ptr(x.device)), "avgpool_fwd")
        ctx.cfg = (n, c, h, w, k)
        return y

    @staticmethod
    def backward(ctx, dy):
        n, c, h, w, k = ctx.cfg
        dy = dy.contiguous(memory_format=torch.channels_last)
        dx = torch.empty((n, c, h, w), dtype=dy.dtype, device=dy.device, memory_format=torch.channels_last)
        nat.check(nat.require().dlb_avgpool_nhwc(1, nat.dtype_code(dy.dtype), dy.data_ptr(), dx.data_ptr(), n, h, w, c, k,
                                                 nat.stream_ptr(dy.device)), "avgpool_bwd")
        return dx, None


def avg_pool2d(x: torch.Tensor, k: int) -> torch.Tensor:
    """``F.avg_pool2d(x, k)`` (stride k)."""
    if x.is_cuda and nat.available() and x.dtype in (torch.float32, torch.bfloat16) and x.shape[1] > 1 \
            and x.shape[2] % k == 0 and x.shape[3] % k == 0:
        return _AvgPoolFn.apply(x, k)
    return F.avg_pool2d(x, k)


def global_avg_pool2d(x: torch.Tensor) -> torch.Tensor:
    """``F.adaptive_avg_pool2d(x, 1)``: the k = H = W case of the average-pool kernel."""
    if x.shape[2] == x.shape[3]:
        return avg_pool2d(x, x.shape[2])
    return F.adaptive_avg_pool2d(x, (1, 1))


class _MaxPoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, k, s, p):
        import ctypes
        lib = nat.require()
        if not getattr(_MaxPoolFn, "_decl", False):
            vp, i32 = ctypes.c_void_p, ctypes.c_int
            nat.declare("dlb_maxpool_nhwc", i32, [i32, i32, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, vp])
            _MaxPoolFn._decl = True
        n, c, h, w = x.shape
        x = x.contiguous(memory_format=torch.channels_last)
        ho, wo = (h + 2 * p - k) // s + 1, (w + 2 * p - k) // s + 1
        y = torch.empty((n, c, ho, wo), dtype=x.dtype, device=x.device, memory_format=torch.channels_last)
        idx = torch.empty((n, ho, wo, c), dtype=torch.uint8, device=x.device)
        nat.check(lib.dlb_maxpool_nhwc(0, nat.dtype_code(x.dtype), x.data_ptr(), y.data_ptr(), idx.data_ptr(), n, h, w, c, k, s, p,
                                       nat.stream_ptr(x.device)), "maxpool_fwd")
        ctx.save_for_backward(idx)
        ctx.cfg = (n, c, h, w, k, s, p)
        return y

    @staticmethod
    def backward(ctx, dy):
        (idx,) = ctx.saved_tensors
        n, c, h, w, k, s, p = ctx.cfg
        dy = dy.contiguous(memory_format=torch.channels_last)
        dx = torch.empty((n, c, h, w), dtype=dy.dtype, device=dy.device, memory_format=torch.channels_last)
        nat.check(nat.require().dlb_maxpool_nhwc(1, nat.dtype_code(dy.dtype), dy.data_ptr(), dx.data_ptr(), idx.data_ptr(), n, h, w, c,
                                                 k, s, p, nat.stream_ptr(dy.device)), "maxpool_bwd")
        return dx, None, None, None


def max_pool2d(x: torch.Tensor, k: int, stride: int = None, padding: int = 0) -> torch.Tensor:
    """``F.max_pool2d(x, k, stride, padding)`` on the NHWC kernel (gather backward, no atomics)."""
    s = k if stride is None else stride
    lib = nat.get()
    if (x.is_cuda and nat.available() and lib is not None and hasattr(lib, "dlb_maxpool_nhwc") and x.dim() == 4
            and x.dtype in (torch.float32, torch.bfloat16) and 2 * padding <= k):
        return _MaxPoolFn.apply(x, k, s, padding)
    return F.max_pool2d(x, k, s, padding)
