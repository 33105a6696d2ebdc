"""LayerNorm(x + residual) as one fused op with a hand-written backward (``csrc/lm.cu``)."""
from __future__ import annotations

import ctypes

import torch

from . import _native as nat

_DECL = False


def _lib():
    global _DECL
    lib = nat.require()
    if not _DECL:
        vp, i32 = ctypes.c_void_p, ctypes.c_int
        nat.declare("dlb_add_layer_norm_fwd", i32, [i32, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, ctypes.c_float, vp])
        nat.declare("dlb_add_layer_norm_bwd", i32, [i32, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, vp])
        _DECL = True
    return lib


class _AddLNFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, residual, weight, bias, eps):
        lib = _lib()
        d = x.shape[-1]
        xc, rc = x.contiguous(), residual.contiguous()
        rows = xc.numel() // d
        y = torch.empty_like(xc)
        z = torch.empty_like(xc)
        mean = torch.empty(rows, dtype=torch.float32, device=x.device)
        rstd = torch.empty(rows, dtype=torch.float32, device=x.device)
        w32 = weight if weight.dtype == torch.float32 else weight.float()
        b32 = bias if bias.dtype == torch.float32 else bias.float()
        nat.check(lib.dlb_add_layer_norm_fwd(nat.dtype_code(xc.dtype), xc.data_ptr(), rc.data_ptr(), w32.data_ptr(), b32.data_ptr(),
                                             y.data_ptr(), z.data_ptr(), mean.data_ptr(), rstd.data_ptr(), rows, d, float(eps),
                                             nat.stream_ptr(x.device)), "add_layer_norm_fwd")
        ctx.save_for_backward(z, mean, rstd, w32)
        ctx.dts = (weight.dtype, bias.dtype)
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _lib()
        z, mean, rstd, w32 = ctx.saved_tensors
        d = z.shape[-1]
        rows = z.numel() // d
        dyc = dy.contiguous()
        dz = torch.empty_like(z)
        dg = torch.empty(d, dtype=torch.float32, device=z.device)
        db = torch.empty(d, dtype=torch.float32, device=z.device)
        nat.check(lib.dlb_add_layer_norm_bwd(nat.dtype_code(z.dtype), dyc.data_ptr(), z.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                             w32.data_ptr(), dz.data_ptr(), dg.data_ptr(), db.data_ptr(), rows, d,
                                             nat.stream_ptr(z.device)), "add_layer_norm_bwd")
        return dz, dz, dg.to(ctx.dts[0]), db.to(ctx.dts[1]), None


def add_layer_norm(x, residual, weight, bias, eps: float = 1e-5):
    return _AddLNFn.apply(x, residual, weight, bias, eps)
