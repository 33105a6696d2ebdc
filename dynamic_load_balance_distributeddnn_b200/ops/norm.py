"""GroupNorm (+ReLU, +residual) as one fused op.

``group_norm_act(x, G, weight, bias, relu=True, residual=None)`` computes
``act(GroupNorm(x) [+ residual])`` — the conv→GN→ReLU / GN→ReLU→conv / GN+shortcut→ReLU
patterns of every CNN in the zoo (reference ``Net/Densenet.py:18-19``, ``Net/Resnet.py:50-54``,
``Net/RegNet.py:57-63``, ``Net/GoogleNet.py:11-37``).

CUDA: hand-written NHWC kernels (``csrc/norm.cu``): 2 launches forward (per-(n,c) stats table +
finalize are fused into one entry point, then one apply pass) and 3 backward, instead of
GN + ReLU (+ add) ATen launches and their separate backward kernels.  CPU / no native library:
the plain PyTorch composition (also the numerics reference in the tests).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn.functional as F

from . import _native as nat


def group_norm_act_reference(x, num_groups, weight, bias, eps=1e-5, relu=True, residual=None):
    y = F.group_norm(x.float(), num_groups, weight.float(), bias.float(), eps)
    if residual is not None:
        y = y + residual.float()
    if relu:
        y = F.relu(y)
    return y.to(x.dtype)


def _nhwc_view(x: torch.Tensor):
    """Return ``(tensor, N, HW, C, row_stride)`` for a logical-NCHW tensor whose memory is NHWC:
    pixel row ``r = (n*H + h)*W + w`` starts at ``base + r*row_stride`` and its C channels are
    contiguous.  Channel slices of a wider channels-last buffer qualify (row_stride > C).  Anything
    else is copied to channels_last first."""
    if x.dim() != 4:
        raise ValueError("expected an NCHW-shaped tensor")
    n, c, h, w = x.shape

    def probe(t):
        sn, sc, sh, sw = t.stride()
        ld = sw if w > 1 else (sh if h > 1 else (sn if n > 1 else c))
        ok = (sc == 1 or c == 1) and ld >= c and (w == 1 or sw == ld) and (h == 1 or sh == w * ld) \
            and (n == 1 or sn == h * w * ld)
        return ok, ld
    ok, ld = probe(x)
    if not ok:
        x = x.contiguous(memory_format=torch.channels_last)
        ok, ld = probe(x)
        if not ok:                       # degenerate shapes: fall back to an explicit NHWC buffer
            x = x.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
            ld = c
    return x, n, h * w, c, ld


class _GNActFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, residual, num_groups, eps, relu, nc_table=None):
        lib = nat.require()
        x, n, hw, c, ldx = _nhwc_view(x)
        y = torch.empty_like(x, memory_format=torch.channels_last)
        _, _, _, _, ldy = _nhwc_view(y)
        res_p, ldr = 0, 0
        if residual is not None:
            residual, _, _, _, ldr = _nhwc_view(residual)
            res_p = residual.data_ptr()
        g = num_groups
        mean = torch.empty(n * g, dtype=torch.float32, device=x.device)
        rstd = torch.empty(n * g, dtype=torch.float32, device=x.device)
        w32 = weight if weight.dtype == torch.float32 else weight.float()
        b32 = bias if bias.dtype == torch.float32 else bias.float()
        st = nat.stream_ptr(x.device)
        if nc_table is not None and nc_table.shape[0] == n and nc_table.shape[1] == c:
            # per-(sample, channel) sums were produced upstream (dense block): no stats pass over x
            nat.check(lib.dlb_gn_finalize(nc_table.data_ptr(), 2 * c, mean.data_ptr(), rstd.data_ptr(), n, c, g, hw,
                                          float(eps), st), "gn_finalize")
            table_ptr, ready = 0, 1
        else:
            table = torch.empty(n * c * 2, dtype=torch.float32, device=x.device)
            table_ptr, ready = table.data_ptr(), 0
        nat.check(lib.dlb_gn_forward(nat.dtype_code(x.dtype), x.data_ptr(), ldx, res_p, ldr, y.data_ptr(), ldy,
                                     w32.data_ptr(), b32.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                     table_ptr, n, hw, c, g, float(eps), int(relu), ready, st), "gn_forward")
        ctx.save_for_backward(x, y if relu else None, w32, mean, rstd)
        ctx.cfg = (n, hw, c, g, ldx, ldy, bool(relu), residual is not None, weight.dtype, bias.dtype)
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = nat.require()
        x, y, w32, mean, rstd = ctx.saved_tensors
        n, hw, c, g, ldx, ldy, relu, has_res, wdt, bdt = ctx.cfg
        dy, _, _, _, lddy = _nhwc_view(dy)
        dx = torch.empty_like(x, memory_format=torch.channels_last)
        _, _, _, _, lddx = _nhwc_view(dx)
        dres = torch.empty_like(x, memory_format=torch.channels_last) if has_res else None
        lddr = _nhwc_view(dres)[4] if has_res else 0
        table = torch.empty(n * c * 2, dtype=torch.float32, device=x.device)
        dgamma = torch.empty(c, dtype=torch.float32, device=x.device)
        dbeta = torch.empty(c, dtype=torch.float32, device=x.device)
        nat.check(lib.dlb_gn_backward(nat.dtype_code(x.dtype), x.data_ptr(), ldx, dy.data_ptr(), lddy,
                                      nat.ptr(y), ldy, dx.data_ptr(), lddx, nat.ptr(dres), lddr,
                                      w32.data_ptr(), mean.data_ptr(), rstd.data_ptr(), table.data_ptr(),
                                      dgamma.data_ptr(), dbeta.data_ptr(), n, hw, c, g, int(relu), 0,
                                      nat.stream_ptr(x.device)), "gn_backward")
        return dx, dgamma.to(wdt), dbeta.to(bdt), dres, None, None, None, None


def group_norm_act(x: torch.Tensor, num_groups: int, weight: torch.Tensor, bias: torch.Tensor,
                   eps: float = 1e-5, relu: bool = True, residual: Optional[torch.Tensor] = None) -> torch.Tensor:
    if x.is_cuda and nat.available() and x.dtype in (torch.float32, torch.bfloat16) and x.dim() == 4:
        return _GNActFn.apply(x, weight, bias, residual, num_groups, eps, relu, getattr(x, "_dlb_nc_table", None))
    return group_norm_act_reference(x, num_groups, weight, bias, eps, relu, residual)
