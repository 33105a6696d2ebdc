"""Stem convolution (3x3 / s1 / p1, <= 4 input channels) on a direct SIMT kernel (``csrc/stem.cu``): the RGB ``conv1`` of every
CNN in the zoo (reference ``Net/Densenet.py:42,76``, ``Net/Resnet.py:63``, ``Net/RegNet.py:70``, ``Net/GoogleNet.py:59``).
Forward + weight gradient; the network input needs no data gradient."""
from __future__ import annotations

import ctypes
import os

import torch

from . import _native as nat
from .norm import _nhwc_view

_DECL = False
ENABLED = os.environ.get("DLB_STEM_CONV", "1") == "1"


def _lib():
    global _DECL
    lib = nat.require()
    if not _DECL:
        vp, i32, i64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong
        nat.declare("dlb_stem_conv_fwd", i32, [i32, vp, vp, vp, i64, i32, i32, i32, i32, i32, vp])
        nat.declare("dlb_stem_conv_wgrad", i32, [i32, vp, vp, i64, vp, i32, i32, i32, i32, i32, vp])
        _DECL = True
    return lib


def supported(x, weight, stride, padding, groups) -> bool:
    lib = nat.get()
    if not (ENABLED and x.is_cuda and nat.available() and lib is not None and hasattr(lib, "dlb_stem_conv_fwd")):
        return False
    if x.dim() != 4 or weight.dim() != 4 or x.requires_grad or x.dtype not in (torch.float32, torch.bfloat16) or weight.dtype != x.dtype:
        return False
    co, ci, kh, kw = weight.shape
    return (kh == 3 and kw == 3 and stride == 1 and padding == 1 and groups == 1 and ci <= 4 and co in (32, 64, 128)
            and x.shape[1] == ci)


class _StemConvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight):
        lib = _lib()
        n, ci, h, w = x.shape
        co = weight.shape[0]
        xv = x.permute(0, 2, 3, 1)
        if not xv.is_contiguous():
            xv = xv.contiguous()                                  # dense NHWC (pixel stride = Cin)
        wk = weight if weight.is_contiguous(memory_format=torch.channels_last) else weight.contiguous(memory_format=torch.channels_last)
        y = torch.empty((n, co, h, w), dtype=x.dtype, device=x.device, memory_format=torch.channels_last)
        nat.check(lib.dlb_stem_conv_fwd(nat.dtype_code(x.dtype), xv.data_ptr(), wk.data_ptr(), y.data_ptr(), co, n, h, w, ci, co,
                                        nat.stream_ptr(x.device)), "stem_conv_fwd")
        ctx.save_for_backward(xv)
        ctx.cfg = (n, h, w, ci, co, weight.dtype)
        ctx.sink = getattr(weight, "_dlb_sink", None)
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _lib()
        (xv,) = ctx.saved_tensors
        n, h, w, ci, co, wdt = ctx.cfg
        dyv, _, _, _, lddy = _nhwc_view(dy)
        flat = ctx.sink.flat() if ctx.sink is not None else None
        if flat is not None and flat.sinks_enabled:
            dw = flat.sink_view(ctx.sink.index)                   # fp32 [Co][3][3][Cin] slice of the flat gradient buffer
            sunk = True
        else:
            dw = torch.zeros((co, 3, 3, ci), dtype=torch.float32, device=dy.device)
            sunk = False
        nat.check(lib.dlb_stem_conv_wgrad(nat.dtype_code(dy.dtype), xv.data_ptr(), dyv.data_ptr(), lddy, dw.data_ptr(), n, h, w, ci, co,
                                          nat.stream_ptr(dy.device)), "stem_conv_wgrad")
        if sunk:
            flat.mark_sunk([ctx.sink.index])
            return None, None
        return None, dw.permute(0, 3, 1, 2).to(wdt)


def conv2d(x, weight):
    return _StemConvFn.apply(x, weight)
