"""Flat parameter / gradient / momentum storage and the fused optimizer step.

The reference keeps 362 separate parameter tensors (DenseNet-121), scales and all-reduces each one
separately and lets ``torch.optim.SGD`` loop over them (``dbs.py:291-301,369,238``).  Here:

* all parameters live in ONE fp32 master buffer (+ one bf16 *shadow* buffer when computing in bf16);
  ``param.data`` of every module parameter is a view into master (fp32-kept params such as norm
  affines) or shadow (bf16 weights) — the model code is unchanged;
* after backward the scattered ``.grad`` tensors are packed by a multi-tensor kernel straight into the
  symmetric allreduce input buffer, applying the rank's DBS weight ``w_r = b_r / B`` and the local
  gradient-clip coefficient on the way (no separate scale / clip kernels);
* the buckets are reduced by the comm back-end into the output buffer;
* one fused kernel applies momentum-SGD to the master buffer and refreshes the bf16 shadow.

A pure-PyTorch implementation of the same steps runs on CPU (gloo debug mode) and is the numerics
reference for the kernels.
"""
from __future__ import annotations

import ctypes
import math
import os
import weakref
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from ..ops import _native as nat
from .comm import Comm

_ALIGN = 32          # elements; keeps every parameter 64-byte aligned in bf16 and 128-byte in fp32


def _is_conv_weight(p: torch.Tensor) -> bool:
    return p.dim() == 4


class GradSink:
    """Handle attached to a parameter (``p._dlb_sink``): ops with hand-written backward kernels write that parameter's
    gradient straight into its slice of the flat (symmetric) gradient buffer instead of returning a ``.grad`` tensor."""
    __slots__ = ("flat", "index")

    def __init__(self, flat: "FlatState", index: int):
        self.flat = weakref.ref(flat)
        self.index = index


def _flat_order(t: torch.Tensor) -> torch.Tensor:
    """1-D view/copy of ``t`` in the order it is stored in the flat buffers: conv weights are kept
    O-H-W-I (channels-last), i.e. exactly the K-major [Cout][kh][kw][Cin] operand layout the implicit-GEMM
    kernels (ours and the vendor's) consume, so no per-call weight re-layout kernel ever runs."""
    if _is_conv_weight(t):
        return t.permute(0, 2, 3, 1).reshape(-1)
    return t.reshape(-1)


class FlatState:
    def __init__(self, model: nn.Module, device, compute_dtype: torch.dtype, comm: Comm,
                 lr: float, momentum: float = 0.9, weight_decay: float = 0.0, bucket_mb: float = 8.0,
                 wire_dtype: torch.dtype = torch.float32, clip_norm: float = 0.0, clip_mode: str = "local",
                 seed_weighting: bool = False):
        self.device = torch.device(device)
        self.comm = comm
        self.params: List[nn.Parameter] = [p for p in model.parameters()]
        self.compute_dtype = compute_dtype
        self.momentum, self.weight_decay = momentum, weight_decay
        # local  = clip this rank's gradient before weighting/allreduce (reference dbs.py:274; coefficient folded into pack)
        # global = clip the REDUCED gradient (coefficient folded into the SGD kernel)
        assert clip_mode in ("local", "global"), clip_mode
        self.clip_norm = clip_norm if clip_mode == "local" else 0.0
        self.global_clip = clip_norm if clip_mode == "global" else 0.0
        self.native = self.device.type == "cuda" and nat.available()
        self.offsets: List[int] = []
        off = 0
        for p in self.params:
            self.offsets.append(off)
            n_alloc = p.numel()
            if p.dim() == 2 and p.shape[0] % 8 != 0:
                # reserve zero rows up to a multiple of 8 behind the matrix (e.g. the 33 278-row vocabulary projection):
                # kernels may then treat it as an aligned [rows_pad, cols] operand; the pad rows receive zero gradients
                rows_pad = (p.shape[0] + 7) // 8 * 8
                n_alloc = rows_pad * p.shape[1]
                p._dlb_padded_rows = rows_pad
            off += (n_alloc + _ALIGN - 1) // _ALIGN * _ALIGN
        self.base_numel = off            # world-size independent part (the tail padding below depends on the world size)
        world = max(1, comm.world)
        pad_to = _ALIGN * world
        self.numel = (off + pad_to - 1) // pad_to * pad_to
        self.master = torch.zeros(self.numel, dtype=torch.float32, device=self.device)
        self.shadow = torch.zeros(self.numel, dtype=torch.bfloat16, device=self.device) \
            if compute_dtype == torch.bfloat16 else None
        self.mom = torch.zeros(self.numel, dtype=torch.float32, device=self.device)
        self.wire_dtype = wire_dtype
        self.grad_in, self.grad_out = comm.alloc_grad_buffers(self.numel, wire_dtype, self.device)
        self.lr_t = torch.full((1,), float(lr), dtype=torch.float32, device=self.device)
        self.weights_t = torch.full((world,), 1.0 / world, dtype=torch.float32, device=self.device)
        self.sumsq_t = torch.zeros(1, dtype=torch.float32, device=self.device)
        self._adopt(model)
        self.buckets = self._make_buckets(bucket_mb)
        self._pack_cache = None
        self._rank = getattr(comm, "rank", 0)
        self._keep_overlap = []
        # ---- gradient path without a scale / pack pass (north star: "no separate elementwise kernel on that path") --------
        # seed weighting: the DBS weight w_r multiplies the loss-gradient seed (fused CE kernel), so every gradient the
        #   backward pass produces is already weighted and the collective runs unweighted (NVLS in-switch reduction);
        # sinks: backward kernels (dense-block wgrad, GroupNorm affine grads) accumulate straight into the flat symmetric
        #   buffer; only the few remaining .grad tensors still go through the multi-tensor pack;
        # fused SGD: the optimizer step runs as the last phase of each bucket's allreduce kernel.
        self.seed_weighting = bool(seed_weighting) and self.native
        self.sinks_enabled = (self.seed_weighting and wire_dtype == torch.float32 and self.clip_norm == 0
                              and os.environ.get("DLB_GRAD_SINKS", "1") == "1")
        self.fused_sgd = (self.native and max(1, comm.world) > 1 and self.global_clip == 0 and hasattr(comm, "allreduce_buckets_sgd")
                          and os.environ.get("DLB_FUSED_SGD", "1") == "1")
        self._sunk = [False] * len(self.params)
        if self.sinks_enabled:
            for i, p in enumerate(self.params):
                p._dlb_sink = GradSink(self, i)

    # ---- parameter adoption -----------------------------------------------------------------------
    def _adopt(self, model: nn.Module) -> None:
        with torch.no_grad():
            for p, off in zip(self.params, self.offsets):
                n = p.numel()
                self.master[off:off + n].copy_(_flat_order(p.detach()).to(self.device, torch.float32))
            if self.shadow is not None:
                self.shadow.copy_(self.master)
            for p, off in zip(self.params, self.offsets):
                n = p.numel()
                keep32 = getattr(p, "_dlb_keep_fp32", False) or self.shadow is None
                src = self.master if keep32 else self.shadow
                if _is_conv_weight(p):
                    o, i, kh, kw = p.shape
                    p.data = src[off:off + n].view(o, kh, kw, i).permute(0, 3, 1, 2)      # logical OIHW, NHWC memory
                else:
                    p.data = src[off:off + n].view(p.shape)
                p.grad = None
        # non-parameter state (buffers such as the positional table) just moves to the device
        for mod in model.modules():
            for k, b in list(mod._buffers.items()):
                if b is not None:
                    mod._buffers[k] = b.to(self.device)

    def _make_buckets(self, bucket_mb: float) -> List[Tuple[int, int]]:
        """Contiguous (offset, numel) ranges of the flat buffer, ~bucket_mb each, cut at parameter boundaries
        (every offset is a multiple of 32 elements, so each range is a whole number of 16-byte vectors).
        ``self.bucket_params[k]`` lists the parameter indices living in bucket k."""
        esize = 2 if self.wire_dtype == torch.bfloat16 else 4
        target = max(_ALIGN, int(bucket_mb * (1 << 20) / esize))
        out, groups = [], []
        start, cur = 0, []
        for i, off in enumerate(self.offsets):
            end = self.offsets[i + 1] if i + 1 < len(self.offsets) else self.numel
            cur.append(i)
            if end - start >= target or i + 1 == len(self.offsets):
                out.append((start, end - start))
                groups.append(cur)
                start, cur = end, []
        if not out:
            out, groups = [(0, self.numel)], [list(range(len(self.params)))]
        self.bucket_params = groups
        return out

    # ---- overlap of the gradient collective with backward -----------------------------------------------
    def enable_overlap(self) -> bool:
        """Fire each bucket's pack + fused allreduce on a side stream as soon as autograd has produced the last
        gradient of the bucket, so the collective overlaps the rest of the backward pass (the reference waits for
        the whole backward and then blocks on one allreduce per tensor, dbs.py:291-301).  Needs the native path,
        more than one rank and no *local* clipping (a global norm is only known after the full backward)."""
        if not self.native or self.comm.world <= 1 or self.clip_norm > 0 or getattr(self, "_overlap", False):
            return getattr(self, "_overlap", False)
        self._overlap = True
        self._comm_stream = torch.cuda.Stream(self.device)
        self._bucket_of = [0] * len(self.params)
        for b, idxs in enumerate(self.bucket_params):
            for i in idxs:
                self._bucket_of[i] = b
        self._pending = [len(g) for g in self.bucket_params]
        self._fired = [False] * len(self.buckets)
        self._hooks = []
        for i, p in enumerate(self.params):
            self._hooks.append(p.register_post_accumulate_grad_hook(lambda _p, i=i: self._on_grad(i)))
        return True

    # ---- gradient sinks ----------------------------------------------------------------------------------
    def sinks_active(self) -> bool:
        return self.sinks_enabled and torch.is_grad_enabled() is not None

    def sink_view(self, i: int) -> torch.Tensor:
        """fp32 view of parameter i's slice of the (zero-initialised, cleared by the optimizer step) gradient buffer, in the
        flat layout: conv weights [O][kh][kw][I], everything else the parameter's own shape."""
        p, off = self.params[i], self.offsets[i]
        v = self.grad_in[off:off + p.numel()]
        if _is_conv_weight(p):
            o, c, kh, kw = p.shape
            return v.view(o, kh, kw, c)
        return v.view(p.shape)

    def mark_sunk(self, indices: Sequence[int]) -> None:
        """The gradients of these parameters have been written into their sinks (on the current stream)."""
        for i in indices:
            self._sunk[i] = True
        if getattr(self, "_overlap", False):
            for i in indices:
                self._on_grad(i)

    def seed_scale(self) -> Optional[torch.Tensor]:
        """device scalar the loss-gradient seed is multiplied by (this rank's DBS weight), or None"""
        return self.weights_t[self._rank:self._rank + 1] if self.seed_weighting else None

    def _sgd_args(self):
        return {"master": self.master.data_ptr(), "mom": self.mom.data_ptr(), "shadow": nat.ptr(self.shadow),
                "lr": self.lr_t.data_ptr(), "momentum": float(self.momentum), "weight_decay": float(self.weight_decay),
                "zero_in": self.grad_in.data_ptr() if self.sinks_enabled else 0}

    def _allreduce(self, buckets) -> float:
        if self.fused_sgd:
            return self.comm.allreduce_buckets_sgd(self.grad_in, self.grad_out, buckets, self._sgd_args())
        return self.comm.allreduce_buckets(self.grad_in, self.grad_out, buckets)

    def _pack(self, subset, rank: int, use_clip: bool = False) -> None:
        lib = nat.require()
        st = nat.stream_ptr(self.device)
        ptrs, offs, numels, dtypes, n, keep = self._grad_lists(subset)
        if n == 0:
            return keep
        if use_clip:
            nat.check(lib.dlb_zero_f32(self.sumsq_t.data_ptr(), 1, st), "zero")
            nat.check(lib.dlb_mt_sumsq(n, ptrs, numels, dtypes, self.sumsq_t.data_ptr(), st), "mt_sumsq")
        nat.check(lib.dlb_mt_pack(n, ptrs, offs, numels, dtypes, self.grad_in.data_ptr(), nat.dtype_code(self.wire_dtype),
                                  None if self.seed_weighting else self.weights_t.data_ptr(), rank,
                                  self.sumsq_t.data_ptr() if use_clip else None, float(self.clip_norm), st), "mt_pack")
        return keep

    def _on_grad(self, i: int) -> None:
        b = self._bucket_of[i]
        self._pending[b] -= 1
        if self._pending[b] == 0 and not self._fired[b]:
            self._fire(b)

    def _fire(self, b: int) -> None:
        main = torch.cuda.current_stream(self.device)
        ev = torch.cuda.Event()
        ev.record(main)
        self._comm_stream.wait_event(ev)
        idxs = [i for i in self.bucket_params[b] if not self._sunk[i]]
        with torch.cuda.stream(self._comm_stream):
            self._keep_overlap.append(self._pack(idxs, self._rank))
            self._allreduce([self.buckets[b]])
        self._fired[b] = True

    # ---- synchronisation helpers ------------------------------------------------------------------
    def sync_initial_params(self) -> None:
        """Average the initial replicas (reference dbs.py:365-367) — a no-op numerically when all
        ranks share the init seed, but it guarantees bit-identical starting points."""
        if self.comm.world > 1:
            self.comm.average_(self.master)
            self.refresh_shadow()

    def refresh_shadow(self) -> None:
        if self.shadow is not None:
            if self.native:
                nat.check(nat.require().dlb_cast_f32_bf16(self.master.data_ptr(), self.shadow.data_ptr(), self.numel,
                                                          nat.stream_ptr(self.device)), "cast")
            else:
                self.shadow.copy_(self.master)

    def set_lr(self, lr: float) -> None:
        self.lr_t.fill_(float(lr))

    def set_weights(self, weights: Sequence[float]) -> None:
        self.weights_t.copy_(torch.as_tensor(list(weights), dtype=torch.float32))
        self._weights_host = [float(w) for w in weights]

    # ---- the post-backward pipeline ---------------------------------------------------------------
    def _grad_lists(self, subset=None):
        ptrs, offs, numels, dtypes, keep = [], [], [], [], []
        if subset is None:
            subset = [i for i in range(len(self.params)) if not self._sunk[i]]
        it = ((self.params[i], self.offsets[i]) for i in subset)
        for p, off in it:
            g = p.grad
            if g is None:
                g = torch.zeros_like(p)
            if _is_conv_weight(p):
                if not g.is_contiguous(memory_format=torch.channels_last):
                    g = g.contiguous(memory_format=torch.channels_last)
            elif not g.is_contiguous():
                g = g.contiguous()
            keep.append(g)
            ptrs.append(g.data_ptr()); offs.append(off); numels.append(g.numel()); dtypes.append(nat.dtype_code(g.dtype))
        n = len(ptrs)
        return ((ctypes.c_void_p * n)(*ptrs), (ctypes.c_longlong * n)(*offs), (ctypes.c_int * n)(*numels),
                (ctypes.c_int * n)(*dtypes), n, keep)

    def reduce_and_step(self, rank: int) -> float:
        """pack(+weight,+clip) → allreduce buckets → fused SGD.  Returns host-measured wait seconds
        (non-zero only for host-blocking back-ends)."""
        if self.native:
            return self._reduce_and_step_native(rank)
        return self._reduce_and_step_torch(rank)

    def _sgd_native(self, g: torch.Tensor) -> None:
        lib = nat.require()
        st = nat.stream_ptr(self.device)
        if g.dtype != torch.float32:
            g = g.float()
        sumsq = None
        if self.global_clip > 0:
            # global L2 norm of the reduced gradient: one multi-tensor sum-of-squares launch over the flat buffer
            n = self.numel
            nat.check(lib.dlb_zero_f32(self.sumsq_t.data_ptr(), 1, st), "zero")
            chunk = 1 << 30
            for o in range(0, n, chunk):
                c = min(chunk, n - o)
                nat.check(lib.dlb_mt_sumsq(1, (ctypes.c_void_p * 1)(g.data_ptr() + 4 * o), (ctypes.c_int * 1)(c),
                                           (ctypes.c_int * 1)(nat.F32), self.sumsq_t.data_ptr(), st), "mt_sumsq")
            sumsq = self.sumsq_t
        nat.check(lib.dlb_sgd_flat_clip(self.master.data_ptr(), self.mom.data_ptr(), g.data_ptr(), nat.ptr(self.shadow), self.numel,
                                        self.lr_t.data_ptr(), float(self.momentum), float(self.weight_decay), nat.ptr(sumsq),
                                        float(self.global_clip), self.grad_in.data_ptr() if self.sinks_enabled else None, st), "sgd_flat")

    def _reduce_and_step_native(self, rank: int) -> float:
        if getattr(self, "_overlap", False):
            # buckets were (mostly) fired from the autograd hooks; flush stragglers, join the comm stream, update
            for b in range(len(self.buckets)):
                if not self._fired[b]:
                    self._fire(b)
            ev = torch.cuda.Event()
            ev.record(self._comm_stream)
            torch.cuda.current_stream(self.device).wait_event(ev)
            if not self.fused_sgd:
                self._sgd_native(self.grad_out)
            self._pending = [len(gp) for gp in self.bucket_params]
            self._fired = [False] * len(self.buckets)
            self._keep, self._keep_overlap = self._keep_overlap, []
            self._sunk = [False] * len(self.params)
            return 0.0
        self._keep = self._pack(None, rank, use_clip=self.clip_norm > 0)
        waited = 0.0
        if self.comm.world > 1:
            waited = self._allreduce(self.buckets)
            if not self.fused_sgd:
                self._sgd_native(self.grad_out)
        else:                              # single rank: the packed (weighted, clipped) gradient IS the result
            self._sgd_native(self.grad_in)
        self._sunk = [False] * len(self.params)
        return waited

    def _reduce_and_step_torch(self, rank: int) -> float:
        with torch.no_grad():
            scale = float(self.weights_t[rank].item())
            if self.clip_norm > 0:
                ss = sum(float((p.grad.float() ** 2).sum()) for p in self.params if p.grad is not None)
                scale *= min(1.0, self.clip_norm / (math.sqrt(ss) + 1e-6))
            self.grad_in.zero_()
            for p, off in zip(self.params, self.offsets):
                if p.grad is not None:
                    self.grad_in[off:off + p.numel()].copy_(_flat_order(p.grad).to(self.grad_in.dtype) * scale)
            waited = self.comm.allreduce_buckets(self.grad_in, self.grad_out, self.buckets)
            g = self.grad_out.float()
            if self.global_clip > 0:
                g = g * min(1.0, self.global_clip / (float(g.norm()) + 1e-6))
            if self.weight_decay:
                g = g + self.weight_decay * self.master
            self.mom.mul_(self.momentum).add_(g)
            self.master.add_(self.mom * (-self.lr_t))
            if self.shadow is not None:
                self.shadow.copy_(self.master)
        return waited

    def zero_grad(self) -> None:
        for p in self.params:
            p.grad = None

    # ---- checkpointing ------------------------------------------------------------------------------
    def state_dict(self):
        n = self.base_numel
        return {"master": self.master[:n].detach().cpu(), "momentum": self.mom[:n].detach().cpu(), "lr": float(self.lr_t.item())}

    def load_state_dict(self, sd) -> None:
        n = min(self.base_numel, sd["master"].numel())        # older checkpoints carry the (zero) tail padding too
        self.master[:n].copy_(sd["master"][:n].to(self.device))
        self.mom[:n].copy_(sd["momentum"][:n].to(self.device))
        self.set_lr(sd["lr"])
        self.refresh_shadow()
