"""Python face of the tcgen05/TMA GEMM (``csrc/gemm_tc.cu``): plain GEMM, 1x1 convolution autograd op,
and the fused GroupNorm-prologue / statistics-epilogue variants used by the dense block."""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import torch

from . import _native as nat
from .norm import _nhwc_view

_DECLARED = False
ENABLED = os.environ.get("DLB_TC_GEMM", "1") == "1"


def _lib():
    global _DECLARED
    lib = nat.require()
    if not _DECLARED:
        vp, i64, i32 = ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int
        nat.declare("dlb_gemm_tc_dt", i32, [i32, vp, i64, vp, i64, vp, i64, i32, i32, i32, vp, vp, i64, i32, vp, i64, i32, vp])
        nat.declare("dlb_gemm_tc_bmn_dt", i32, [i32, vp, i64, vp, i64, vp, i64, i32, i32, i32, i32, vp])
        nat.declare("dlb_conv3x3_tc_dt", i32, [i32, i32, vp, i64, vp, vp, i64, i32, i32, i32, i32, i32, vp, i64, i32, vp])
        nat.declare("dlb_wgrad3x3_tc_dt", i32, [i32, vp, i64, vp, i64, vp, i32, i32, i32, i32, i32, i32, vp])
        nat.declare("dlb_wgrad_tc_dt", i32, [i32, vp, i64, vp, i64, vp, i64, i32, i32, i32, vp, vp, i64, i32, i32, vp])
        nat.declare("dlb_dgrad_gn_dt", i32, [i32, i32, vp, i64, vp, i64, vp, i64, vp, i64, i32, i32, i32, i32, vp, vp, vp, vp, i64, vp, i64,
                                             i32, vp])
        nat.declare("dlb_gn_bwd_coeff", i32, [vp, i64, vp, vp, vp, vp, vp, i64, vp, vp, i32, i32, i32, i32, vp])
        _DECLARED = True
    return lib


def available() -> bool:
    return ENABLED and nat.available() and hasattr(nat.get(), "dlb_gemm_tc_dt")


#: element types the tensor-core kernels take: bf16 (kind::f16) and fp32 storage with TF32 math (kind::tf32) -- the
#: precision class of the reference's default PyTorch path (fp32 tensors, cuDNN convolutions with allow_tf32=True)
TC_DTYPES = (torch.bfloat16, torch.float32) if os.environ.get("DLB_TF32", "1") == "1" else (torch.bfloat16,)
BF16, F32 = nat.BF16, nat.F32


def vec(dtype_code: int) -> int:
    """elements per 16-byte vector: every K / N / row stride must be a multiple of it"""
    return 4 if dtype_code == F32 else 8


def gemm_raw(a_ptr: int, lda: int, b_ptr: int, ldb: int, d_ptr: int, ldd: int, m: int, n: int, k: int, device,
             pro_a: Optional[torch.Tensor] = None, pro_b: Optional[torch.Tensor] = None, rows_per_sample: int = 0,
             stats: Optional[torch.Tensor] = None, stats_ptr: int = 0, stats_ns: int = 0, sm_limit: int = 0,
             dtype: int = nat.BF16) -> None:
    pro_ld = pro_a.shape[1] if pro_a is not None else 0
    sp = stats_ptr if stats_ptr else nat.ptr(stats)
    rc = _lib().dlb_gemm_tc_dt(dtype, a_ptr, lda, b_ptr, ldb, d_ptr, ldd, m, n, k, nat.ptr(pro_a), nat.ptr(pro_b), pro_ld,
                            rows_per_sample, sp, stats_ns, sm_limit, nat.stream_ptr(device))
    nat.check(rc, "gemm_tc")


def gemm(a: torch.Tensor, b: torch.Tensor, out: Optional[torch.Tensor] = None, pro_a=None, pro_b=None,
         rows_per_sample: int = 0, stats=None, stats_ns: int = 0) -> torch.Tensor:
    """out[M,N] = pro(a[M,K]) @ b[N,K]^T.  a, b, out: bf16 or fp32 (TF32 math) 2-D with unit inner stride (row strides free)."""
    assert a.dtype in TC_DTYPES and b.dtype == a.dtype and a.stride(1) == 1 and b.stride(1) == 1
    m, k = a.shape
    n = b.shape[0]
    if out is None:
        out = torch.empty((m, n), dtype=a.dtype, device=a.device)
    assert out.stride(1) == 1
    gemm_raw(a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0), out.data_ptr(), out.stride(0), m, n, k, a.device,
             pro_a, pro_b, rows_per_sample, stats, 0, stats_ns, dtype=nat.dtype_code(a.dtype))
    return out


def gemm_bmn_raw(a_ptr: int, lda: int, b_ptr: int, ldb: int, d_ptr: int, ldd: int, m: int, n: int, k: int, device,
                 sm_limit: int = 0, dtype: int = nat.BF16) -> None:
    """d[m,n] = a[m,k] @ b[k,n] with b row-major [k][n] (no transposed copy; MN-major tensor-core operand)."""
    nat.check(_lib().dlb_gemm_tc_bmn_dt(dtype, a_ptr, lda, b_ptr, ldb, d_ptr, ldd, m, n, k, sm_limit, nat.stream_ptr(device)),
              "gemm_tc_bmn")


def gemm_bmn(a: torch.Tensor, b_kn: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    m, k = a.shape
    n = b_kn.shape[1]
    if out is None:
        out = torch.empty((m, n), dtype=a.dtype, device=a.device)
    gemm_bmn_raw(a.data_ptr(), a.stride(0), b_kn.data_ptr(), b_kn.stride(0), out.data_ptr(), out.stride(0), m, n, k, a.device,
                 dtype=nat.dtype_code(a.dtype))
    return out


def wgrad_raw(dy_ptr: int, lddy: int, x_ptr: int, ldx: int, dw: torch.Tensor, m: int, co: int, ci: int, device,
              pro_a: Optional[torch.Tensor] = None, pro_b: Optional[torch.Tensor] = None, rows_per_sample: int = 0,
              sm_limit: int = 0, dtype: int = nat.BF16) -> None:
    """dw[co, ci] (fp32, pre-zeroed or accumulated into) += dy[m, co]^T @ pro(x[m, ci])."""
    pro_ld = pro_a.shape[1] if pro_a is not None else 0
    rc = _lib().dlb_wgrad_tc_dt(dtype, dy_ptr, lddy, x_ptr, ldx, dw.data_ptr(), dw.stride(0), m, co, ci, nat.ptr(pro_a), nat.ptr(pro_b),
                             pro_ld, rows_per_sample, sm_limit, nat.stream_ptr(device))
    nat.check(rc, "wgrad_tc")


def wgrad(dy: torch.Tensor, x: torch.Tensor, pro_a=None, pro_b=None, rows_per_sample: int = 0) -> torch.Tensor:
    """-> fp32 [Co, Ci] = dy[M,Co]^T @ pro(x[M,Ci]); dy/x bf16 or fp32 2-D with unit inner stride."""
    m, co = dy.shape
    ci = x.shape[1]
    dw = torch.zeros((co, ci), dtype=torch.float32, device=dy.device)
    wgrad_raw(dy.data_ptr(), dy.stride(0), x.data_ptr(), x.stride(0), dw, m, co, ci, dy.device, pro_a, pro_b, rows_per_sample,
              dtype=nat.dtype_code(dy.dtype))
    return dw


# ---- dgrad GEMM fused with the GroupNorm(+ReLU) backward of its input, csrc/dgrad_gn.cu (DLB_FUSED_DGRAD=0 disables) ----
# validated on B200 in round 2: kernel tests vs fp64, stage-level gradients no further from the fp32 truth than the
# three-kernel chain (tests/test_gpu_dgrad_gn.py), -5 % step time at batch 512 and -6 % at batch 64 (bf16)
FUSED_DGRAD = os.environ.get("DLB_FUSED_DGRAD", "1") == "1"
# the fp32-storage / TF32-math flavour of the same kernels (DLB_FUSED_DGRAD_TF32=0 keeps tf32 on the chain)
FUSED_DGRAD_TF32 = os.environ.get("DLB_FUSED_DGRAD_TF32", "1") == "1"


def dgrad_gn_available() -> bool:
    return available() and hasattr(nat.get(), "dlb_dgrad_gn_dt")


def fused_dgrad_enabled(dtype: torch.dtype) -> bool:
    if not (FUSED_DGRAD and dgrad_gn_available()):
        return False
    return dtype == torch.bfloat16 or (dtype == torch.float32 and FUSED_DGRAD_TF32)


def dgrad_gn_raw(mode: int, dy_ptr: int, lddy: int, w_ptr: int, ldw: int, x_ptr: int, ldx: int, dx_ptr: int, lddx: int,
                 m: int, n: int, k: int, rows_per_sample: int, ca: torch.Tensor, cb: torch.Tensor,
                 k2: Optional[torch.Tensor], k3: Optional[torch.Tensor], table_ptr: int, table_ns: int, device,
                 sm_limit: int = 0, dtype: int = nat.BF16) -> None:
    """mode 1: table += per-(sample, channel) (sum dz, sum dz*x) with dz = (dy @ w) * [ca*x + cb > 0];
    mode 2: dx += ca*dz + k2*x + k3 in place.  dy [m,k], w [k,n] row-major, x/dx [m,n] with free row strides."""
    rc = _lib().dlb_dgrad_gn_dt(dtype, mode, dy_ptr, lddy, w_ptr, ldw, x_ptr, ldx, dx_ptr, lddx, m, n, k, rows_per_sample, ca.data_ptr(),
                             cb.data_ptr(), nat.ptr(k2), nat.ptr(k3), ca.shape[1], table_ptr, table_ns, sm_limit,
                             nat.stream_ptr(device))
    nat.check(rc, f"dgrad_gn(mode {mode})")


def gn_bwd_coeff_raw(table_ptr: int, table_ns: int, gamma: torch.Tensor, mean: torch.Tensor, rstd: torch.Tensor, k2: torch.Tensor,
                     k3: torch.Tensor, dgamma_ptr: int, dbeta_ptr: int, n: int, c: int, groups: int, hw: int, device) -> None:
    rc = _lib().dlb_gn_bwd_coeff(table_ptr, table_ns, gamma.data_ptr(), mean.data_ptr(), rstd.data_ptr(), k2.data_ptr(),
                                 k3.data_ptr(), k2.shape[1], dgamma_ptr, dbeta_ptr, n, c, groups, hw, nat.stream_ptr(device))
    nat.check(rc, "gn_bwd_coeff")


CONV3_MAX_HW = int(os.environ.get("DLB_TC_CONV3_MAX_HW", "64"))


def conv3x3_geometry_ok(h: int, w: int) -> bool:
    if w <= 0 or 128 % w:
        return False
    rows = 128 // w
    return (h % rows == 0) if rows <= h else (rows % h == 0)


HALO = os.environ.get("DLB_CONV3_HALO", "1") == "1"


def conv3x3_halo_ok(h: int, w: int) -> bool:
    """shapes the halo flavour of the 3x3 kernel takes (csrc/gemm_tc.cu, HALO): 128-pixel tiles made of whole rows of ONE image,
    (rows + 2) * W <= 192 pixels per stage, row offsets in whole 1024-byte swizzle groups -- i.e. 16x16 and 32x32 maps"""
    if not HALO or w <= 0 or 128 % w or w % 8:
        return False
    rows = 128 // w
    return rows <= h and h % rows == 0 and (rows + 2) * w <= 192


def conv3x3_profitable(h: int, w: int) -> bool:
    """Measured on B200 (tools/bench_gemm.py, 128->32 channels): the tap-by-tap implicit GEMM re-reads its input tile
    nine times from L2; it wins or ties against the vendor kernel at 8x8 / 4x4 feature maps and lost 1.3-1.9x at
    16x16 / 32x32 -- those maps take the halo flavour instead (input rows loaded once per horizontal tap: 3 L2->SM passes
    instead of 9).  Anything else above DLB_TC_CONV3_MAX_HW pixels goes to the vendor library."""
    return conv3x3_geometry_ok(h, w) and (h * w <= CONV3_MAX_HW or conv3x3_halo_ok(h, w))


def conv3x3_raw(dgrad: bool, x_ptr: int, ldx: int, w_ptr: int, y_ptr: int, ldy: int, n: int, h: int, w: int, ci: int, co: int,
                device, stats_ptr: int = 0, stats_ns: int = 0, sm_limit: int = 0, dtype: int = nat.BF16) -> None:
    """3x3/s1/p1 NHWC conv on tcgen05: forward (x[N,H,W,ci] -> y[N,H,W,co]) or data gradient (x := dY[..,co] -> y := dX[..,ci]).
    Weight memory must be [co][3][3][ci] (channels-last OIHW).  ldx / ldy are pixel strides in elements."""
    rc = _lib().dlb_conv3x3_tc_dt(dtype, int(dgrad), x_ptr, ldx, w_ptr, y_ptr, ldy, n, h, w, ci, co, stats_ptr, stats_ns, sm_limit,
                               nat.stream_ptr(device))
    nat.check(rc, "conv3x3_tc")


WGRAD3 = os.environ.get("DLB_TC_WGRAD3", "1") == "1"


def wgrad3x3_supported(n: int, h: int, w: int, ci: int, dtype: torch.dtype) -> bool:
    return (WGRAD3 and available() and hasattr(nat.get(), "dlb_wgrad3x3_tc_dt") and dtype in TC_DTYPES and conv3x3_geometry_ok(h, w)
            and ci % (8 if dtype == torch.bfloat16 else 4) == 0)


def wgrad3x3_raw(x_ptr: int, ldx: int, dy_ptr: int, lddy: int, dw: torch.Tensor, n: int, h: int, w: int, ci: int, co: int, device,
                 dtype: int = nat.BF16, sm_limit: int = 0) -> None:
    """dw[co][3][3][ci] (fp32, contiguous, pre-zeroed or accumulated into) += 3x3/s1/p1 weight gradient; x [n,h,w,ci] and
    dy [n,h,w,co] are NHWC with pixel strides ldx / lddy (channel slices of wider buffers are read in place)."""
    assert dw.dtype == torch.float32 and dw.is_contiguous() and dw.numel() == co * 9 * ci
    rc = _lib().dlb_wgrad3x3_tc_dt(dtype, x_ptr, ldx, dy_ptr, lddy, dw.data_ptr(), n, h, w, ci, co, sm_limit, nat.stream_ptr(device))
    nat.check(rc, "wgrad3x3_tc")


def _w_ohwi(weight: torch.Tensor) -> torch.Tensor:
    """weight [O,I,3,3] -> tensor whose memory is [O][3][3][I] (no copy when already channels-last)."""
    return weight if weight.is_contiguous(memory_format=torch.channels_last) else weight.contiguous(memory_format=torch.channels_last)


class _Conv3x3Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight):
        xv, n, hw, c, ld = _nhwc_view(x)
        o = weight.shape[0]
        h, w = x.shape[2], x.shape[3]
        wk = _w_ohwi(weight)
        y = torch.empty((n, o, h, w), dtype=x.dtype, device=x.device, memory_format=torch.channels_last)
        conv3x3_raw(False, xv.data_ptr(), ld, wk.data_ptr(), y.data_ptr(), o, n, h, w, c, o, x.device, dtype=nat.dtype_code(x.dtype))
        ctx.save_for_backward(xv, weight)
        ctx.cfg = (n, h, w, c, ld, o)
        return y

    @staticmethod
    def backward(ctx, dy):
        xv, weight = ctx.saved_tensors
        n, h, w, c, ld, o = ctx.cfg
        dyv, _, _, _, lddy = _nhwc_view(dy)
        dx = dw = None
        if ctx.needs_input_grad[0]:
            wk = _w_ohwi(weight)
            dx = torch.empty((n, c, h, w), dtype=dy.dtype, device=dy.device, memory_format=torch.channels_last)
            conv3x3_raw(True, dyv.data_ptr(), lddy, wk.data_ptr(), dx.data_ptr(), c, n, h, w, c, o, dy.device,
                        dtype=nat.dtype_code(dy.dtype))
        if ctx.needs_input_grad[1] and wgrad3x3_supported(n, h, w, c, dy.dtype):
            # 9-tap MN-major split-K tcgen05 weight gradient (csrc/conv_wgrad.cu), both operands read in place
            dwf = torch.zeros((o, 3, 3, c), dtype=torch.float32, device=dy.device)
            wgrad3x3_raw(xv.data_ptr(), ld, dyv.data_ptr(), lddy, dwf, n, h, w, c, o, dy.device, dtype=nat.dtype_code(dy.dtype))
            dw = dwf.permute(0, 3, 1, 2).to(weight.dtype)              # logical OIHW, channels-last memory
        elif ctx.needs_input_grad[1]:
            x4 = torch.as_strided(xv, (n, c, h, w), (h * w * ld, 1, w * ld, ld))
            _, dw, _ = torch.ops.aten.convolution_backward(dyv if dyv.is_contiguous(memory_format=torch.channels_last) else
                                                           dyv.contiguous(memory_format=torch.channels_last),
                                                           x4 if ld == c else x4.contiguous(memory_format=torch.channels_last),
                                                           weight, None, [1, 1], [1, 1], [1, 1], False, [0, 0], 1,
                                                           [False, True, False])
        return dx, dw


def _w2d(weight: torch.Tensor) -> torch.Tensor:
    """[O, I, 1, 1] (any layout) -> contiguous-row [O, I] view."""
    o, i = weight.shape[0], weight.shape[1]
    w = weight.reshape(o, i)
    return w if w.stride(1) == 1 else w.contiguous()


class _Conv1x1Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight):
        xv, n, hw, c, ld = _nhwc_view(x)
        o = weight.shape[0]
        h, w = x.shape[2], x.shape[3]
        y = torch.empty((n, o, h, w), dtype=x.dtype, device=x.device, memory_format=torch.channels_last)
        w2 = _w2d(weight)
        gemm_raw(xv.data_ptr(), ld, w2.data_ptr(), w2.stride(0), y.data_ptr(), o, n * hw, o, c, x.device, dtype=nat.dtype_code(x.dtype))
        ctx.save_for_backward(xv, weight)
        ctx.cfg = (n, hw, c, ld, o, h, w)
        return y

    @staticmethod
    def backward(ctx, dy):
        xv, weight = ctx.saved_tensors
        n, hw, c, ld, o, h, w = ctx.cfg
        dyv, _, _, _, lddy = _nhwc_view(dy)
        dx = dw = None
        w2 = _w2d(weight)
        if ctx.needs_input_grad[0]:
            dx = torch.empty((n, c, h, w), dtype=dy.dtype, device=dy.device, memory_format=torch.channels_last)
            # dX = dY * W with W = [Cout][Cin] consumed as an MN-major operand: no transposed weight copy
            gemm_bmn_raw(dyv.data_ptr(), lddy, w2.data_ptr(), w2.stride(0), dx.data_ptr(), c, n * hw, c, o, dy.device,
                         dtype=nat.dtype_code(dy.dtype))
        if ctx.needs_input_grad[1]:
            dwf = torch.zeros((o, c), dtype=torch.float32, device=dy.device)
            wgrad_raw(dyv.data_ptr(), lddy, xv.data_ptr(), ld, dwf, n * hw, o, c, dy.device,     # MN-major split-K tcgen05
                      dtype=nat.dtype_code(dy.dtype))
            dw = dwf.view(o, c, 1, 1).to(weight.dtype)
        return dx, dw


def conv_supported(x, weight, stride, padding, groups) -> bool:
    if not available() or x.dtype not in TC_DTYPES or weight.dtype != x.dtype or x.dim() != 4 or groups != 1:
        return False
    v = 8 if x.dtype == torch.bfloat16 else 4
    if weight.dim() == 4 and weight.shape[2] == 3 and weight.shape[3] == 3 and stride == 1 and padding == 1:
        return (x.shape[1] % v == 0 and weight.shape[0] % v == 0 and conv3x3_profitable(x.shape[2], x.shape[3])
                and x.shape[0] * x.shape[2] * x.shape[3] >= 128)
    return (stride == 1 and padding == 0 and weight.shape[2] == 1 and weight.shape[3] == 1
            and x.shape[1] % v == 0 and weight.shape[0] % v == 0 and x.shape[0] * x.shape[2] * x.shape[3] >= 128)


class _LinearFn(torch.autograd.Function):
    """y[T,N] = x[T,K] @ W[N,K]^T on the tcgen05 GEMM; dgrad consumes W as an MN-major operand, wgrad reads dY and X
    as MN-major operands (reduction over tokens) -- no transposes anywhere."""

    @staticmethod
    def forward(ctx, x2, weight):
        t, k = x2.shape
        n = weight.shape[0]
        y = torch.empty((t, n), dtype=x2.dtype, device=x2.device)
        gemm_raw(x2.data_ptr(), x2.stride(0), weight.data_ptr(), weight.stride(0), y.data_ptr(), n, t, n, k, x2.device,
                 dtype=nat.dtype_code(x2.dtype))
        ctx.save_for_backward(x2, weight)
        return y

    @staticmethod
    def backward(ctx, dy):
        x2, weight = ctx.saved_tensors
        t, k = x2.shape
        n = weight.shape[0]
        dy = dy.contiguous()
        dx = dw = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty((t, k), dtype=dy.dtype, device=dy.device)
            gemm_bmn_raw(dy.data_ptr(), n, weight.data_ptr(), weight.stride(0), dx.data_ptr(), k, t, k, n, dy.device,
                         dtype=nat.dtype_code(dy.dtype))
        if ctx.needs_input_grad[1]:
            dwf = torch.zeros((n, k), dtype=torch.float32, device=dy.device)
            wgrad_raw(dy.data_ptr(), n, x2.data_ptr(), x2.stride(0), dwf, t, n, k, dy.device, dtype=nat.dtype_code(dy.dtype))
            dw = dwf.to(weight.dtype)
        return dx, dw


class _LinearPadFn(torch.autograd.Function):
    """Small classifier heads (N = 10 / 100 outputs: not a multiple of the 16-byte vector) on the same tcgen05 GEMMs: the flat
    parameter store keeps zero rows behind such a matrix up to a multiple of 8 (``_dlb_padded_rows``), so the weight is
    consumed as an aligned [Np, K] operand, the output / its gradient live in [T, Np] buffers and the caller sees the
    first N columns (reference classifier heads: Net/Densenet.py:83, Net/Resnet.py:87, Net/RegNet.py:104)."""

    @staticmethod
    def forward(ctx, x2, weight, n_pad):
        t, k = x2.shape
        n = weight.shape[0]
        wp = torch.as_strided(weight, (n_pad, k), (weight.stride(0), 1))
        y = torch.empty((t, n_pad), dtype=x2.dtype, device=x2.device)
        gemm_raw(x2.data_ptr(), x2.stride(0), wp.data_ptr(), wp.stride(0), y.data_ptr(), n_pad, t, n_pad, k, x2.device,
                 dtype=nat.dtype_code(x2.dtype))
        ctx.save_for_backward(x2, wp)
        ctx.n = n
        return y[:, :n]

    @staticmethod
    def backward(ctx, dy):
        x2, wp = ctx.saved_tensors
        t, k = x2.shape
        n, n_pad = ctx.n, wp.shape[0]
        dyp = torch.zeros((t, n_pad), dtype=x2.dtype, device=x2.device)
        dyp[:, :n].copy_(dy)
        dx = dw = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty((t, k), dtype=x2.dtype, device=x2.device)
            gemm_bmn_raw(dyp.data_ptr(), n_pad, wp.data_ptr(), wp.stride(0), dx.data_ptr(), k, t, k, n_pad, x2.device,
                         dtype=nat.dtype_code(x2.dtype))
        if ctx.needs_input_grad[1]:
            dwf = torch.zeros((n_pad, k), dtype=torch.float32, device=x2.device)
            wgrad_raw(dyp.data_ptr(), n_pad, x2.data_ptr(), x2.stride(0), dwf, t, n_pad, k, x2.device, dtype=nat.dtype_code(x2.dtype))
            dw = dwf[:n].to(wp.dtype)
        return dx, dw, None


def linear_supported(x, weight) -> bool:
    if not available() or x.dtype not in TC_DTYPES or weight.dtype != x.dtype or weight.dim() != 2:
        return False
    n, k = weight.shape
    v = 8 if x.dtype == torch.bfloat16 else 4
    tokens = x.numel() // max(1, x.shape[-1])
    if n % v != 0:
        # small heads: only through the zero-padded rows of the flat parameter store (see _LinearPadFn)
        n_pad = int(getattr(weight, "_dlb_padded_rows", 0))
        return (n_pad >= n and n_pad % v == 0 and k % v == 0 and tokens >= 32 and weight.stride(1) == 1 and weight.stride(0) == k
                and weight.is_contiguous())
    return k % v == 0 and tokens >= 128 and weight.stride(1) == 1 and weight.stride(0) % v == 0


def linear(x, weight, bias=None):
    k = x.shape[-1]
    x2 = x.reshape(-1, k)
    if x2.stride(1) != 1 or x2.stride(0) % (8 if x2.dtype == torch.bfloat16 else 4) != 0 or (x2.data_ptr() & 15):
        x2 = x2.contiguous()
    n = weight.shape[0]
    if n % (8 if x2.dtype == torch.bfloat16 else 4) != 0:
        y = _LinearPadFn.apply(x2, weight, int(weight._dlb_padded_rows))            # [T, N] view of a [T, Np] buffer
        if bias is not None:
            y = y + bias.to(y.dtype)
        return y.reshape(*x.shape[:-1], n)
    y = _LinearFn.apply(x2, weight).view(*x.shape[:-1], n)
    if bias is not None:
        y = y + bias.to(y.dtype)
    return y


def conv2d(x, weight, bias, stride, padding):
    y = _Conv3x3Fn.apply(x, weight) if weight.shape[2] == 3 else _Conv1x1Fn.apply(x, weight)
    if bias is not None:
        y = y + bias.to(y.dtype).view(1, -1, 1, 1)
    return y
