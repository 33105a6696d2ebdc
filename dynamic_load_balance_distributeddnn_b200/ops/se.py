"""Squeeze-and-excitation as one fused kernel per direction (``csrc/se.cu``; reference ``Net/RegNet.py:10-25``)."""
from __future__ import annotations

import ctypes

import torch
import torch.nn.functional as F

from . import _native as nat
from .norm import _nhwc_view

_DECL = False


def _lib():
    global _DECL
    lib = nat.require()
    if not _DECL:
        vp, i32, i64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong
        nat.declare("dlb_se_fwd", i32, [i32, vp, i64, vp, i64, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, vp])
        nat.declare("dlb_se_bwd", i32, [i32, vp, i64, vp, i64, vp, i64, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, vp])
        _DECL = True
    return lib


def supported(x, w1, b1, w2, b2) -> bool:
    lib = nat.get()
    return (x.is_cuda and nat.available() and lib is not None and hasattr(lib, "dlb_se_fwd") and x.dim() == 4
            and x.dtype in (torch.float32, torch.bfloat16) and all(t is not None and t.dtype == x.dtype for t in (w1, b1, w2, b2))
            and x.shape[1] <= 1024 and w1.shape[0] <= 256)


class _SEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2):
        lib = _lib()
        xv, n, hw, c, ld = _nhwc_view(x)
        cs = w1.shape[0]
        w1m, w2m = w1.reshape(cs, c), w2.reshape(c, cs)
        w1m = w1m if w1m.is_contiguous() else w1m.contiguous()
        w2m = w2m if w2m.is_contiguous() else w2m.contiguous()
        out = torch.empty((n, c, x.shape[2], x.shape[3]), dtype=x.dtype, device=x.device, memory_format=torch.channels_last)
        s = torch.empty((n, c), dtype=torch.float32, device=x.device)
        h = torch.empty((n, cs), dtype=torch.float32, device=x.device)
        z = torch.empty((n, c), dtype=torch.float32, device=x.device)
        nat.check(lib.dlb_se_fwd(nat.dtype_code(x.dtype), xv.data_ptr(), ld, out.data_ptr(), c, w1m.data_ptr(), b1.data_ptr(), w2m.data_ptr(),
                                 b2.data_ptr(), s.data_ptr(), h.data_ptr(), z.data_ptr(), n, hw, c, cs, nat.stream_ptr(x.device)), "se_fwd")
        ctx.save_for_backward(xv, w1m, w2m, s, h, z)
        ctx.cfg = (n, hw, c, ld, cs, x.shape[2], x.shape[3], w1.shape, w2.shape)
        return out

    @staticmethod
    def backward(ctx, dout):
        lib = _lib()
        xv, w1m, w2m, s, h, z = ctx.saved_tensors
        n, hw, c, ld, cs, hh, ww, w1s, w2s = ctx.cfg
        dv, _, _, _, ldg = _nhwc_view(dout)
        dx = torch.empty((n, c, hh, ww), dtype=dout.dtype, device=dout.device, memory_format=torch.channels_last)
        dpre1 = torch.empty((n, cs), dtype=torch.float32, device=dout.device)
        dpre2 = torch.empty((n, c), dtype=torch.float32, device=dout.device)
        nat.check(lib.dlb_se_bwd(nat.dtype_code(dout.dtype), xv.data_ptr(), ld, dv.data_ptr(), ldg, dx.data_ptr(), c, w1m.data_ptr(),
                                 w2m.data_ptr(), h.data_ptr(), z.data_ptr(), dpre1.data_ptr(), dpre2.data_ptr(), n, hw, c, cs,
                                 nat.stream_ptr(dout.device)), "se_bwd")
        wdt = w1m.dtype
        dw1 = (dpre1.t() @ s).to(wdt).view(w1s)              # [cs, c]   tiny matrix products over the batch
        dw2 = (dpre2.t() @ h).to(wdt).view(w2s)              # [c, cs]
        return dx, dw1, dpre1.sum(0).to(wdt), dw2, dpre2.sum(0).to(wdt)


def squeeze_excite(x, w1, b1, w2, b2):
    """x * sigmoid(W2 relu(W1 mean_hw(x) + b1) + b2); w1 [CS, C, 1, 1], w2 [C, CS, 1, 1]"""
    if supported(x, w1, b1, w2, b2):
        return _SEFn.apply(x, w1, b1, w2, b2)
    s = F.adaptive_avg_pool2d(x, (1, 1))
    s = F.conv2d(F.relu(F.conv2d(s, w1.to(x.dtype), b1.to(x.dtype))), w2.to(x.dtype), b2.to(x.dtype)).sigmoid()
    return x * s
