"""Fused token-embedding front-end: ``dropout(E[tokens] * sqrt(d) + pe[:S])`` in one kernel, scatter-add backward
(``csrc/embed.cu``; reference ``Net/Transformer.py:48-49,91-92``)."""
from __future__ import annotations

import ctypes
import math
from typing import Dict

import torch
import torch.nn.functional as F

from . import _native as nat

_DECL = False
_STEP: Dict[str, torch.Tensor] = {}
_SEED = [0x2545F491]


def _lib():
    global _DECL
    lib = nat.require()
    if not _DECL:
        vp, i32, i64, f32, u32 = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_float, ctypes.c_uint
        nat.declare("dlb_embed_fwd", i32, [i32, vp, vp, i64, vp, i64, vp, i32, i32, i32, i32, f32, f32, u32, vp, vp])
        nat.declare("dlb_embed_bwd", i32, [i32, vp, vp, vp, i64, i32, i32, i32, i32, f32, f32, u32, vp, vp])
        _DECL = True
    return lib


def available() -> bool:
    lib = nat.get()
    return nat.available() and lib is not None and hasattr(lib, "dlb_embed_fwd")


def set_seed(seed: int) -> None:
    _SEED[0] = int(seed) & 0xFFFFFFFF


def _step_tensor(device) -> torch.Tensor:
    key = str(device)
    if key not in _STEP:
        _STEP[key] = torch.zeros(1, dtype=torch.int64, device=device)
    return _STEP[key]


class _EmbedFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, tokens, weight, pe, scale, p_drop):
        lib = _lib()
        s, b = tokens.shape
        v, d = weight.shape
        tokens = tokens.contiguous()
        out = torch.empty((s, b, d), dtype=weight.dtype, device=weight.device)
        step = _step_tensor(weight.device)
        step_used = step.clone()                     # the backward regenerates the same dropout mask
        step += 1
        pe2 = pe.reshape(pe.shape[0], -1)
        nat.check(lib.dlb_embed_fwd(nat.dtype_code(weight.dtype), tokens.data_ptr(), weight.data_ptr(), weight.stride(0), pe2.data_ptr(),
                                    pe2.stride(0), out.data_ptr(), s, b, d, v, float(scale), float(p_drop), _SEED[0], step_used.data_ptr(),
                                    nat.stream_ptr(weight.device)), "embed_fwd")
        ctx.save_for_backward(tokens, step_used)
        ctx.cfg = (s, b, d, v, float(scale), float(p_drop), _SEED[0], weight.dtype)
        return out

    @staticmethod
    def backward(ctx, dout):
        lib = _lib()
        tokens, step_used = ctx.saved_tensors
        s, b, d, v, scale, p_drop, seed, wdt = ctx.cfg
        dout = dout.contiguous()
        dw = torch.zeros((v, d), dtype=torch.float32, device=dout.device)
        nat.check(lib.dlb_embed_bwd(nat.dtype_code(dout.dtype), tokens.data_ptr(), dout.data_ptr(), dw.data_ptr(), d, s, b, d, v, scale,
                                    p_drop, seed, step_used.data_ptr(), nat.stream_ptr(dout.device)), "embed_bwd")
        return None, dw.to(wdt), None, None, None


def supported(tokens: torch.Tensor, weight: torch.Tensor) -> bool:
    return (tokens.is_cuda and available() and tokens.dtype == torch.int64 and tokens.dim() == 2 and weight.dim() == 2
            and weight.dtype in (torch.float32, torch.bfloat16) and weight.stride(1) == 1)


def embed_pe_dropout(tokens: torch.Tensor, weight: torch.Tensor, pe: torch.Tensor, p_drop: float, training: bool) -> torch.Tensor:
    """tokens int64 [S, B]; weight [V, d]; pe fp32 [max_len, 1, d] -> [S, B, d]"""
    scale = math.sqrt(weight.shape[1])
    p = p_drop if training else 0.0
    if supported(tokens, weight):
        return _EmbedFn.apply(tokens, weight, pe, scale, p)
    x = F.embedding(tokens, weight) * scale
    return F.dropout(x + pe[:x.size(0)].to(x.dtype), p, training)
