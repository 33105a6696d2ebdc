"""Communication back-ends behind one interface.

``Comm`` is what the trainer talks to; three implementations:

* :class:`SymmComm` — the product.  Gradient buckets live in symmetric memory and each bucket's
  weighted allreduce is ONE hand-written sm_100a kernel (one-shot / two-shot / NVLS, ``csrc/comm.cu``)
  launched on the caller's stream; the per-rank time exchange is a single P2P-store kernel; barriers
  are device-side.  No NCCL call on the data path.
* :class:`TorchComm` with ``nccl`` — the A/B baseline: the same flat buckets through
  ``dist.all_reduce`` (already far better than the reference's per-parameter loop).
* :class:`TorchComm` with ``gloo`` — CPU debug mode (reference ``-d true``; SURVEY §4's "fake
  multi-GPU backend").

The collective set mirrors the reference's five call sites: group init (C1), initial
parameter averaging (C2), barrier (C3), weighted gradient allreduce (C4), per-rank time all-gather (C5).
"""
from __future__ import annotations

import ctypes
import os
import time
from typing import List, Optional, Sequence

import torch
import torch.distributed as dist

from ..ops import _native as nat
from .symm import SymmetricAllocator

ALGO_CODES = {"oneshot": 0, "twoshot": 1, "nvls": 2}


class Comm:
    rank: int = 0
    world: int = 1
    name: str = "none"

    def barrier(self) -> None: ...
    def average_(self, flat: torch.Tensor) -> None: ...
    def broadcast_(self, flat: torch.Tensor, src: int = 0) -> None: ...
    def gather_times(self, value: float) -> List[float]:
        return [float(value)]

    def allreduce_buckets(self, grad_in: torch.Tensor, grad_out: torch.Tensor, buckets: Sequence[tuple]) -> float:
        """Sum ``grad_in`` (already weighted) across ranks into ``grad_out``; returns host-measured wait
        seconds (0 when the wait is measured on the device)."""
        if grad_out.data_ptr() != grad_in.data_ptr():
            grad_out.copy_(grad_in)
        return 0.0

    def alloc_grad_buffers(self, numel: int, dtype: torch.dtype, device):
        g = torch.zeros(numel, dtype=dtype, device=device)
        return g, g

    def device_wait_seconds(self, reset: bool = True) -> float:
        return 0.0

    def check_errors(self) -> None: ...

    def close(self) -> None: ...


class SingleComm(Comm):
    name = "single"


class TorchComm(Comm):
    """Flat-bucket allreduce through torch.distributed (gloo on CPU, nccl on GPU)."""

    def __init__(self, group=None):
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.name = dist.get_backend(group)

    def barrier(self) -> None:
        dist.barrier(group=self.group)

    def average_(self, flat: torch.Tensor) -> None:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
        flat.div_(float(self.world))

    def broadcast_(self, flat: torch.Tensor, src: int = 0) -> None:
        dist.broadcast(flat, src=src, group=self.group)

    def gather_times(self, value: float) -> List[float]:
        dev = "cuda" if self.name == "nccl" else "cpu"
        t = torch.tensor([float(value)], dtype=torch.float32, device=dev)
        out = [torch.empty_like(t) for _ in range(self.world)]
        dist.all_gather(out, t, group=self.group)
        return [float(o.item()) for o in out]

    def allreduce_buckets(self, grad_in, grad_out, buckets) -> float:
        t0 = time.perf_counter()
        works = [dist.all_reduce(grad_in[o:o + n], op=dist.ReduceOp.SUM, group=self.group, async_op=True)
                 for (o, n) in buckets]
        for w in works:
            w.wait()
        waited = time.perf_counter() - t0 if not grad_in.is_cuda else 0.0
        if grad_out.data_ptr() != grad_in.data_ptr():
            grad_out.copy_(grad_in)
        return waited


class SymmComm(Comm):
    """Fused weighted allreduce over symmetric memory (the north-star path)."""

    name = "symm"

    def __init__(self, device, group=None, algo: str = "auto", blocks: int = 0, timeout_s: float = 20.0,
                 backend: str = "auto"):
        self.device = torch.device(device)
        self.group = group
        self.lib = nat.require()
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.alloc = SymmetricAllocator(self.device, group, backend)
        self.algo = algo
        self.blocks = blocks
        self.timeout_s = timeout_s
        # one-warp gate kernel in front of every bucket's allreduce (csrc/comm.cu: dlb_comm_gate); DLB_COMM_GATE=0 disables
        self.gate = os.environ.get("DLB_COMM_GATE", "1") == "1" and hasattr(self.lib, "dlb_comm_gate")
        self._ctx = None
        self._time_ctx = None
        self._counters = torch.zeros(4, dtype=torch.int64, device=self.device)   # [0]=wait_ns, [1]=err flag (int32 view)
        self._err_view = self._counters[1:2].view(torch.int32)[:1]
        flag_bytes = int(self.lib.dlb_comm_flag_words()) * 4
        self._flags = self.alloc.alloc("flags", flag_bytes)
        self._times = self.alloc.alloc("times", 256)
        self._my_time = torch.zeros(1, dtype=torch.float32, device=self.device)
        self._make_time_ctx()
        self._sync_host()

    # ---- plumbing -------------------------------------------------------------------------------
    def _sync_host(self) -> None:
        torch.cuda.synchronize(self.device)
        if self.world > 1:
            dist.barrier(group=self.group)

    def _ptr_array(self, ptrs):
        return (ctypes.c_ulonglong * len(ptrs))(*ptrs)

    def _make_ctx(self, in_buf, out_buf):
        ctx = self.lib.dlb_comm_create(self.rank, self.world, self._ptr_array(in_buf.ptrs), self._ptr_array(out_buf.ptrs),
                                       self._ptr_array(self._flags.ptrs), in_buf.multicast_ptr, out_buf.multicast_ptr,
                                       self._counters.data_ptr(), self._err_view.data_ptr())
        if not ctx:
            raise RuntimeError("dlb_comm_create failed")
        self.lib.dlb_comm_set_timeout(ctx, float(self.timeout_s))
        return ctx

    def _make_time_ctx(self):
        self._time_ctx = self._make_ctx(self._times, self._times)

    def alloc_grad_buffers(self, numel: int, dtype: torch.dtype, device):
        esize = torch.empty(0, dtype=dtype).element_size()
        self._in = self.alloc.alloc("grad_in", numel * esize)
        self._out = self.alloc.alloc("grad_out", numel * esize)
        self._wire = dtype
        self._ctx = self._make_ctx(self._in, self._out)
        self._sync_host()
        return self._in.view(dtype, numel), self._out.view(dtype, numel)

    @property
    def has_multicast(self) -> bool:
        return self._ctx is not None and self._in.multicast_ptr != 0 and self._out.multicast_ptr != 0

    # ---- collectives ----------------------------------------------------------------------------
    def pick_algo(self, nbytes: int) -> str:
        """Measured on 8xB200 (profiles/r1_11): with the DBS weight already applied by the pack kernel the NVLS variant
        (multimem.ld_reduce / multimem.st, no staging pass) is the fastest from 1 MiB up (809 GB/s busbw at 256 MiB vs
        598 for scale+NCCL); two-shot P2P is next (and the choice without a multicast mapping); one-shot only wins the
        latency race for tiny buckets."""
        if self.algo != "auto":
            if self.algo == "nvls" and not self.has_multicast:
                return "twoshot"
            return self.algo
        if self.world == 1 or nbytes <= 32 * 1024:
            return "oneshot"
        # in-switch reduction pays off with the number of peers: at 2 GPUs the two-shot P2P kernel is ~1.6x faster than the
        # multimem path (profiles/r2_06_allreduce_sweep_n2.json: 57 vs 105 us at 16 MiB, 586 vs 375 GB/s busbw at 256 MiB)
        if self.has_multicast and nbytes >= 512 * 1024 and self.world > 2:
            return "nvls"
        return "twoshot"

    def pick_blocks(self, nbytes: int, algo: str) -> int:
        if self.blocks > 0:
            return self.blocks
        if nbytes <= 64 * 1024:
            return 4
        if nbytes <= 1 << 20:
            return 16
        if nbytes <= 32 << 20:
            return 32
        return 64

    def allreduce_buckets(self, grad_in, grad_out, buckets, weights_dev: Optional[torch.Tensor] = None) -> float:
        st = nat.stream_ptr(self.device)
        wire = nat.dtype_code(self._wire)
        esize = 2 if self._wire == torch.bfloat16 else 4
        for (off, n) in buckets:
            algo = self.pick_algo(n * esize)
            if self.gate and self.world > 1:
                nat.check(self.lib.dlb_comm_gate(self._ctx, st), "comm_gate")
            rc = self.lib.dlb_weighted_allreduce(self._ctx, ALGO_CODES[algo], wire, off, n, self.pick_blocks(n * esize, algo),
                                                 nat.ptr(weights_dev), None, st)
            nat.check(rc, "weighted_allreduce")
        return 0.0

    def allreduce_buckets_sgd(self, grad_in, grad_out, buckets, sgd: dict) -> float:
        """Allreduce of each bucket with the momentum-SGD step (+ bf16 shadow refresh, + clearing of the gradient buffer)
        fused behind it as the kernel's last phase: ONE launch per bucket for "collective + optimizer"."""
        st = nat.stream_ptr(self.device)
        wire = nat.dtype_code(self._wire)
        esize = 2 if self._wire == torch.bfloat16 else 4
        for (off, n) in buckets:
            algo = self.pick_algo(n * esize)
            if self.gate and self.world > 1:
                nat.check(self.lib.dlb_comm_gate(self._ctx, st), "comm_gate")
            rc = self.lib.dlb_weighted_allreduce_sgd(self._ctx, ALGO_CODES[algo], wire, off, n, self.pick_blocks(n * esize, algo),
                                                     None, sgd["master"], sgd["mom"], sgd["shadow"], sgd["lr"], sgd["zero_in"],
                                                     sgd["momentum"], sgd["weight_decay"], None, st)
            nat.check(rc, "weighted_allreduce_sgd")
        return 0.0

    def barrier(self) -> None:
        nat.check(self.lib.dlb_device_barrier(self._time_ctx, 2, nat.stream_ptr(self.device)), "device_barrier")

    def average_(self, flat: torch.Tensor) -> None:
        # one-shot use of the fused allreduce with uniform weights (reference dbs.py:365-367)
        if self.world == 1:
            return
        gin = self._in.view(flat.dtype) if flat.dtype == self._wire else None
        if gin is None or gin.numel() < flat.numel():
            dist.all_reduce(flat, group=self.group)
            flat.div_(self.world)
            return
        n = flat.numel()
        n_al = (n + 31) // 32 * 32
        gin[:n].copy_(flat)
        w = torch.full((self.world,), 1.0 / self.world, dtype=torch.float32, device=self.device)
        self.allreduce_buckets(None, None, [(0, n_al)], w)
        flat.copy_(self._out.view(flat.dtype)[:n])
        self.check_errors()

    def broadcast_(self, flat: torch.Tensor, src: int = 0) -> None:
        if self.world > 1:
            dist.broadcast(flat, src=src, group=self.group)

    def gather_times(self, value: float) -> List[float]:
        if self.world == 1:
            return [float(value)]
        self._my_time.fill_(float(value))
        par = getattr(self, "_time_parity", 0)
        self._time_parity = par ^ 1                      # double-buffered table: a fast peer's next
        nat.check(self.lib.dlb_time_allgather(self._time_ctx, self._my_time.data_ptr(), par * 16,   # store cannot clobber
                                              nat.stream_ptr(self.device)), "time_allgather")
        table = self._times.view(torch.float32, 32)[par * 16: par * 16 + self.world]
        out = table.cpu().tolist()           # stream-ordered D2H: sees the kernel's result
        self.check_errors()
        return [float(v) for v in out]

    def device_wait_seconds(self, reset: bool = True) -> float:
        ns = int(self._counters[0].item())
        if reset:
            self._counters[0].zero_()
        return ns * 1e-9

    def check_errors(self) -> None:
        e = int(self._err_view.item())
        if e != 0:
            raise RuntimeError(f"device collective watchdog fired (code {e}): a peer did not arrive within "
                               f"{self.timeout_s}s")

    def close(self) -> None:
        for c in (self._ctx, self._time_ctx):
            if c:
                self.lib.dlb_comm_destroy(c)
        self._ctx = self._time_ctx = None


def make_comm(kind: str, device, group=None, **kw) -> Comm:
    """kind: symm | nccl | gloo | single."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        if kind == "symm" and torch.device(device).type == "cuda":
            return SymmComm(device, group, **kw)
        return SingleComm()
    if kind == "symm":
        return SymmComm(device, group, **kw)
    return TorchComm(group)
