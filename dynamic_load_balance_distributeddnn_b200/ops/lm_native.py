"""Native (sm_100a) fast paths for the language-model ops; each ``has_*`` reports whether the
kernel is present in the built library so ``transformer_ops`` can fall back explicitly."""
from __future__ import annotations

import ctypes

import torch

from . import _native as nat

_DECL = False


def _lib():
    global _DECL
    lib = nat.require()
    if not _DECL:
        vp, i32, i64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong
        nat.declare("dlb_softmax_ce_inplace", i32, [i32, vp, i64, vp, vp, vp, i32, i32, ctypes.c_float, vp])
        _DECL = True
    return lib


def has_add_layer_norm() -> bool:
    lib = nat.get()
    return lib is not None and hasattr(lib, "dlb_add_layer_norm_fwd")


def has_linear_ce() -> bool:
    lib = nat.get()
    return lib is not None and hasattr(lib, "dlb_softmax_ce_inplace")


def add_layer_norm(x, residual, weight, bias, eps):
    from .layernorm import add_layer_norm as f
    return f(x, residual, weight, bias, eps)


class _FusedLinearCE(torch.autograd.Function):
    """loss = mean CE(feats @ W^T + b, target).  180 GB of HBM means the [T, V] logits (1.2 GB in bf16 for the
    wikitext-2 step) can simply exist once: one GEMM writes them, ONE kernel turns them into the loss and, in place,
    into d(loss)/d(logits); the backward is two more GEMMs and a column sum.  No chunk loop, no fp32 logits."""

    @staticmethod
    def forward(ctx, feats, weight, bias, target):
        lib = _lib()
        from . import gemm_tc
        t, d = feats.shape
        v = weight.shape[0]
        w = weight if weight.dtype == feats.dtype else weight.to(feats.dtype)
        vp = int(getattr(weight, "_dlb_padded_rows", 0))
        tc = (vp >= v and vp % 8 == 0 and w is weight and feats.dtype == torch.bfloat16 and weight.is_contiguous()
              and gemm_tc.available() and d % 8 == 0 and t >= 128)
        if tc:
            # the flat parameter store keeps zero rows behind the vocabulary matrix up to a multiple of 8, so all three
            # GEMMs of the loss run on the tcgen05 kernels with an aligned [Vp, d] operand
            wp = torch.as_strided(weight, (vp, d), (d, 1))
            logits = gemm_tc.gemm(feats, wp)                                         # [T, Vp]
        else:
            wp = None
            logits = feats @ w.t()                                                   # [T, V]
        loss = torch.zeros(1, dtype=torch.float32, device=feats.device)
        b32 = None if bias is None else (bias if bias.dtype == torch.float32 else bias.float())
        nat.check(lib.dlb_softmax_ce_inplace(nat.dtype_code(logits.dtype), logits.data_ptr(), logits.stride(0), nat.ptr(b32),
                                             target.data_ptr(), loss.data_ptr(), t, v, 1.0 / t, nat.stream_ptr(feats.device)),
                  "softmax_ce_inplace")
        ctx.save_for_backward(logits, feats, w if wp is None else wp)               # `logits` now holds d loss / d logits
        ctx.cfg = (weight.dtype, None if bias is None else bias.dtype, tc, v)
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        from . import gemm_tc
        dlogits, feats, w = ctx.saved_tensors
        wdt, bdt, tc, v = ctx.cfg
        if tc:
            dfeats = gemm_tc.gemm_bmn(dlogits, w) * g.to(dlogits.dtype)               # [T,Vp] x [Vp,d]
            dweight = (gemm_tc.wgrad(dlogits, feats)[:v] * g.float()).to(wdt)        # fp32 [Vp,d] -> [V,d]
            dbias = None if bdt is None else (dlogits[:, :v].sum(0, dtype=torch.float32) * g.float()).to(bdt)
            return dfeats, dweight, dbias, None
        g = g.to(dlogits.dtype)
        dfeats = (dlogits @ w) * g
        dweight = ((dlogits.t() @ feats) * g).to(wdt)
        dbias = None if bdt is None else (dlogits.sum(0, dtype=torch.float32) * g.float()).to(bdt)
        return dfeats, dweight, dbias, None


def linear_cross_entropy(feats, weight, bias, target):
    return _FusedLinearCE.apply(feats.contiguous(), weight, bias, target.contiguous())
