"""Fused causal self-attention for short sequences (``csrc/lm.cu``): reads the packed in-projection output
[S, B, 3*D] in place, keeps scores / softmax / dropout / PV on chip, hand-written backward."""
from __future__ import annotations

import ctypes
import math
from typing import Dict

import torch

from . import _native as nat

_DECL = False
_STEP: Dict[str, torch.Tensor] = {}
_SEED = [0x5DEECE66]


def _lib():
    global _DECL
    lib = nat.require()
    if not _DECL:
        vp, i32, f32 = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
        nat.declare("dlb_attention_fwd", i32, [i32, vp, vp, vp, i32, i32, i32, i32, f32, f32, ctypes.c_uint, vp, vp])
        nat.declare("dlb_attention_bwd", i32, [i32, vp, vp, vp, vp, i32, i32, i32, i32, f32, f32, ctypes.c_uint, vp, vp])
        _DECL = True
    return lib


def available() -> bool:
    lib = nat.get()
    return nat.available() and lib is not None and hasattr(lib, "dlb_attention_fwd")


def set_seed(seed: int) -> None:
    _SEED[0] = int(seed) & 0xFFFFFFFF


def _step_tensor(device) -> torch.Tensor:
    key = str(device)
    if key not in _STEP:
        _STEP[key] = torch.zeros(1, dtype=torch.int64, device=device)
    return _STEP[key]


class _AttnFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, qkv, nhead, p_drop):
        lib = _lib()
        s, b, d3 = qkv.shape
        d = d3 // 3
        hd = d // nhead
        qkv = qkv.contiguous()
        out = torch.empty((s, b, d), dtype=qkv.dtype, device=qkv.device)
        probs = torch.empty((b, nhead, s, s), dtype=torch.float32, device=qkv.device)
        step = _step_tensor(qkv.device)
        step_used = step.clone()                   # the backward must regenerate the same dropout mask
        step += 1
        scale = 1.0 / math.sqrt(hd)
        nat.check(lib.dlb_attention_fwd(nat.dtype_code(qkv.dtype), qkv.data_ptr(), out.data_ptr(), probs.data_ptr(), s, b, nhead, hd,
                                        scale, float(p_drop), _SEED[0], step_used.data_ptr(), nat.stream_ptr(qkv.device)), "attention_fwd")
        ctx.save_for_backward(qkv, probs, step_used)
        ctx.cfg = (s, b, nhead, hd, scale, float(p_drop), _SEED[0])
        return out

    @staticmethod
    def backward(ctx, dout):
        lib = _lib()
        qkv, probs, step_used = ctx.saved_tensors
        s, b, nhead, hd, scale, p_drop, seed = ctx.cfg
        dqkv = torch.empty_like(qkv)
        dout = dout.contiguous()
        nat.check(lib.dlb_attention_bwd(nat.dtype_code(qkv.dtype), qkv.data_ptr(), dout.data_ptr(), probs.data_ptr(), dqkv.data_ptr(),
                                        s, b, nhead, hd, scale, p_drop, seed, step_used.data_ptr(), nat.stream_ptr(qkv.device)),
                  "attention_bwd")
        return dqkv, None, None


def causal_attention_packed(qkv: torch.Tensor, nhead: int, p_drop: float) -> torch.Tensor:
    """qkv [S, B, 3*D] -> [S, B, D]"""
    return _AttnFn.apply(qkv, nhead, p_drop)


def supported(qkv: torch.Tensor, nhead: int) -> bool:
    if not (qkv.is_cuda and available() and qkv.dtype in (torch.float32, torch.bfloat16) and qkv.dim() == 3):
        return False
    s, _, d3 = qkv.shape
    return s <= 64 and (d3 // 3) // nhead <= 128
