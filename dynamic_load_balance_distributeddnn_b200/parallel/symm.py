"""Symmetric (peer-mapped) device memory for the fused collective kernels.

Every rank allocates the same set of named buffers; after ``rendezvous`` each rank holds the
device-virtual addresses of *all* ranks' copies, which the kernels in ``csrc/comm.cu`` load from /
store to directly over NVLink (and, when a multicast mapping exists, reduce inside the NVSwitch).

Allocator back-ends, tried in order:
  1. ``torch.distributed._symmetric_memory`` — cuMem VMM allocations exported as POSIX fds, plus an
     NVLS multicast binding when the fabric supports it.  Used purely as an allocator/rendezvous;
     no torch collective runs on the data path.
  2. native CUDA-IPC (``csrc/ipc.cu``) with the 64-byte handles exchanged through
     ``torch.distributed.all_gather_object`` — no multicast, so ``nvls`` is disabled.
  3. world_size == 1: plain tensors.

torch.distributed itself is only the bootstrap (reference ``dbs.py:513-515`` uses it for
everything).
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch
import torch.distributed as dist

from ..ops import _native as nat


@dataclass
class SymmBuffer:
    name: str
    tensor: torch.Tensor                 # local uint8 view
    ptrs: List[int]                      # base address of every rank's copy (index = rank)
    multicast_ptr: int = 0
    backend: str = "local"
    _keep: list = field(default_factory=list)

    def view(self, dtype: torch.dtype, numel: Optional[int] = None, offset_bytes: int = 0) -> torch.Tensor:
        t = self.tensor[offset_bytes:].view(dtype)
        return t if numel is None else t[:numel]


class _CudaArray:
    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


class SymmetricAllocator:
    def __init__(self, device: torch.device, group=None, backend: str = "auto"):
        self.device = torch.device(device)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.backend = backend if backend != "auto" else os.environ.get("DLB_SYMM_BACKEND", "auto")
        self.buffers: Dict[str, SymmBuffer] = {}

    # ------------------------------------------------------------------ back-ends
    def _alloc_local(self, name: str, nbytes: int) -> SymmBuffer:
        t = torch.zeros(nbytes, dtype=torch.uint8, device=self.device)
        return SymmBuffer(name, t, [t.data_ptr()], 0, "local")

    def _alloc_torch_symm(self, name: str, nbytes: int) -> SymmBuffer:
        import torch.distributed._symmetric_memory as symm_mem
        group = self.group if self.group is not None else dist.group.WORLD
        t = symm_mem.empty(nbytes, dtype=torch.uint8, device=self.device)
        try:
            hdl = symm_mem.rendezvous(t, group=group)
        except Exception:
            # older torch needs the group enabled explicitly (deprecated no-op on 2.11)
            symm_mem.enable_symm_mem_for_group(group.group_name)
            hdl = symm_mem.rendezvous(t, group=group)
        t.zero_()
        ptrs = [int(p) for p in hdl.buffer_ptrs]
        try:
            mc = int(getattr(hdl, "multicast_ptr", 0) or 0)
        except Exception:
            mc = 0
        return SymmBuffer(name, t, ptrs, mc, "torch_symm", [hdl])

    def _alloc_ipc(self, name: str, nbytes: int) -> SymmBuffer:
        lib = nat.require()
        for fn, res, args in (("dlb_ipc_alloc", ctypes.c_int, [ctypes.c_ulonglong, ctypes.c_void_p]),
                              ("dlb_ipc_get_handle", ctypes.c_int, [ctypes.c_ulonglong, ctypes.c_void_p]),
                              ("dlb_ipc_open", ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p])):
            nat.declare(fn, res, args)
        ptr = ctypes.c_ulonglong(0)
        nbytes_al = (nbytes + (2 << 20) - 1) // (2 << 20) * (2 << 20)
        nat.check(lib.dlb_ipc_alloc(nbytes_al, ctypes.byref(ptr)), "ipc_alloc")
        handle = (ctypes.c_ubyte * 64)()
        nat.check(lib.dlb_ipc_get_handle(ptr.value, handle), "ipc_get_handle")
        handles: List[Optional[bytes]] = [None] * self.world
        dist.all_gather_object(handles, bytes(handle), group=self.group)
        ptrs = []
        for r, h in enumerate(handles):
            if r == self.rank:
                ptrs.append(int(ptr.value))
            else:
                out = ctypes.c_ulonglong(0)
                buf = (ctypes.c_ubyte * 64).from_buffer_copy(h)
                nat.check(lib.dlb_ipc_open(buf, ctypes.byref(out)), "ipc_open")
                ptrs.append(int(out.value))
        arr = _CudaArray(int(ptr.value), nbytes)
        t = torch.as_tensor(arr, device=self.device)
        return SymmBuffer(name, t, ptrs, 0, "cuda_ipc", [arr])

    # ------------------------------------------------------------------ public
    def alloc(self, name: str, nbytes: int) -> SymmBuffer:
        nbytes = (int(nbytes) + 255) // 256 * 256
        if self.world == 1:
            buf = self._alloc_local(name, nbytes)
        else:
            buf, err = None, None
            order = ["torch_symm", "cuda_ipc"] if self.backend == "auto" else [self.backend]
            for b in order:
                try:
                    buf = self._alloc_torch_symm(name, nbytes) if b == "torch_symm" else self._alloc_ipc(name, nbytes)
                    # every rank must agree on the backend
                    ok = torch.tensor([1], device=self.device)
                    dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
                    if int(ok.item()) == 1:
                        break
                    buf = None
                except Exception as e:            # noqa: BLE001 - try next back-end, report the last error
                    err = e
                    try:
                        ok = torch.tensor([0], device=self.device)
                        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
                    except Exception:
                        pass
                    buf = None
            if buf is None:
                raise RuntimeError(f"no symmetric-memory backend available: {err!r}")
        self.buffers[name] = buf
        return buf
