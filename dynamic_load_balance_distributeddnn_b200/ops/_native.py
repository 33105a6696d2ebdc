"""ctypes binding of ``libdlb_b200.so`` (hand-written sm_100a kernels, C ABI; ``csrc/``).

The library is built in-tree by ``__graft_entry__.build()`` / ``make -C csrc`` and travels to the
GPU box with the snapshot.  On a CUDA machine a missing library is a hard error (no silent
fallback to eager PyTorch); on a CPU-only machine ``available()`` is simply False and the ops use
their plain PyTorch reference implementations.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import c_double, c_float, c_int, c_longlong, c_uint, c_ulonglong, c_void_p
from typing import Optional

import torch

_PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.environ.get("DLB_NATIVE_LIB") or os.path.join(_PKG_DIR, "libdlb_b200.so")   # override: A/B two builds
CSRC_DIR = os.path.join(_PKG_DIR, "csrc")

_lib: Optional[ctypes.CDLL] = None
_tried = False

F32, BF16 = 0, 1


def dtype_code(dt: torch.dtype) -> int:
    if dt == torch.float32:
        return F32
    if dt == torch.bfloat16:
        return BF16
    raise TypeError(f"unsupported dtype {dt}")


def build(verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a into the in-tree shared library."""
    res = subprocess.run(["make", "-C", CSRC_DIR, "-j", str(min(16, os.cpu_count() or 4))],
                         capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout[-4000:])
        print(res.stderr[-4000:])
    if res.returncode != 0:
        raise RuntimeError("native build failed")
    return LIB_PATH


def _declare(lib: ctypes.CDLL) -> None:
    vp, i64, i32 = c_void_p, c_longlong, c_int
    sig = {
        "dlb_launch_count": (c_ulonglong, []),
        "dlb_launch_count_add": (None, [c_ulonglong]),
        "dlb_nc_reduce2": (i32, [i32, i32, vp, i64, vp, i64, vp, i64, vp, i64, i32, i32, i32, vp]),
        "dlb_gn_finalize": (i32, [vp, i64, vp, vp, i32, i32, i32, i32, c_float, vp]),
        "dlb_nc_reduce2_bwd": (i32, [i32, i32, vp, i64, vp, i64, vp, i64, vp, i64, vp, vp, vp, vp, i32, i32, i32, i32, vp]),
        "dlb_nc_reduce2_bwd_coef": (i32, [i32, vp, i64, vp, i64, vp, i64, vp, vp, vp, vp, vp, vp, i64, i32, i32, i32, i32, vp]),
        "dlb_gn_bwd_apply_coef": (i32, [i32, vp, i64, vp, i64, vp, i64, vp, vp, vp, vp, i64, vp, vp, i64, i32, i32, i32, i32, i32, vp]),
        "dlb_gn_fwd_apply_table": (i32, [i32, vp, i64, vp, i64, vp, vp, vp, i64, vp, vp, vp, vp, i64, i32, i32, i32, i32, c_float, i32, vp]),
        "dlb_gn_bwd_fused": (i32, [i32, vp, i64, vp, i64, vp, i64, vp, vp, vp, vp, i64, vp, vp, vp, vp, i64, vp, i32, i32, i32, i32, i32, vp]),
        "dlb_copy_stats": (i32, [i32, vp, i64, vp, i64, vp, i64, i32, i32, i32, vp]),
        "dlb_norm_skip_zero": (None, [i32]),
        "dlb_norm_bulk": (None, [i32, i32]),
        "dlb_gn_coeff": (i32, [vp, i64, vp, vp, vp, vp, vp, vp, i64, i32, i32, i32, i32, c_float, vp]),
        "dlb_gn_fwd_apply": (i32, [i32, vp, i64, vp, i64, vp, i64, vp, vp, vp, vp, i32, i32, i32, i32, i32, vp]),
        "dlb_gn_bwd_apply": (i32, [i32, vp, i64, vp, i64, vp, i64, vp, i64, vp, i64, vp, vp, vp, vp, i64,
                                   i32, i32, i32, i32, i32, i32, vp]),
        "dlb_gn_param_grad": (i32, [vp, i64, vp, vp, vp, vp, i32, i32, i32, vp]),
        "dlb_copy2d": (i32, [vp, i64, vp, i64, i64, i32, vp]),
        "dlb_avgpool_nhwc": (i32, [i32, i32, vp, vp, i32, i32, i32, i32, i32, vp]),
        "dlb_gn_forward": (i32, [i32, vp, i64, vp, i64, vp, i64, vp, vp, vp, vp, vp, i32, i32, i32, i32,
                                 c_float, i32, i32, vp]),
        "dlb_gn_backward": (i32, [i32, vp, i64, vp, i64, vp, i64, vp, i64, vp, i64, vp, vp, vp, vp, vp, vp,
                                  i32, i32, i32, i32, i32, i32, vp]),
        "dlb_mt_sumsq": (i32, [i32, vp, vp, vp, vp, vp]),
        "dlb_mt_pack": (i32, [i32, vp, vp, vp, vp, vp, i32, vp, i32, vp, c_float, vp]),
        "dlb_sgd_flat": (i32, [vp, vp, vp, vp, i64, vp, c_float, c_float, vp]),
        "dlb_sgd_flat_clip": (i32, [vp, vp, vp, vp, i64, vp, c_float, c_float, vp, c_float, vp, vp]),
        "dlb_softmax_ce_small": (i32, [i32, vp, i64, vp, vp, vp, vp, i32, i32, vp]),
        "dlb_weighted_allreduce_sgd": (i32, [vp, i32, i32, i64, i64, i32, vp, vp, vp, vp, vp, vp, c_float, c_float, vp, vp]),
        "dlb_cast_f32_bf16": (i32, [vp, vp, i64, vp]),
        "dlb_zero_f32": (i32, [vp, i64, vp]),
        "dlb_augment": (i32, [vp, vp, i32, i32, i32, i32, i32, i32, i32, vp, vp, c_uint, vp, vp]),
        "dlb_burn": (i32, [vp, c_float, vp]),
        "dlb_stamp": (i32, [vp, vp]),
        "dlb_stamp_acc": (i32, [vp, vp, vp]),
        "dlb_comm_create": (vp, [i32, i32, vp, vp, vp, c_ulonglong, c_ulonglong, c_ulonglong, c_ulonglong]),
        "dlb_comm_destroy": (None, [vp]),
        "dlb_comm_set_timeout": (None, [vp, c_double]),
        "dlb_comm_flag_words": (i32, []),
        "dlb_comm_max_blocks": (i32, []),
        "dlb_weighted_allreduce": (i32, [vp, i32, i32, i64, i64, i32, vp, vp, vp]),
        "dlb_time_allgather": (i32, [vp, vp, i32, vp]),
        "dlb_device_barrier": (i32, [vp, i32, vp]),
        "dlb_comm_gate": (i32, [vp, vp]),
    }
    for name, (res, args) in sig.items():
        if hasattr(lib, name):
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
    # optional symbols (later build stages) are declared by their own modules via `declare()`


def declare(name: str, restype, argtypes) -> bool:
    lib = get()
    if lib is None or not hasattr(lib, name):
        return False
    fn = getattr(lib, name)
    fn.restype = restype
    fn.argtypes = argtypes
    return True


def get() -> Optional[ctypes.CDLL]:
    global _lib, _tried
    if _lib is not None or _tried:
        return _lib
    _tried = True
    if not os.path.isfile(LIB_PATH):
        if torch.cuda.is_available() and os.environ.get("DLB_ALLOW_NO_NATIVE", "0") != "1":
            # be loud on a GPU box: try one in-place build, else fail
            build()
        else:
            return None
    _lib = ctypes.CDLL(LIB_PATH)
    _declare(_lib)
    if hasattr(_lib, "dlb_set_pdl"):
        _lib.dlb_set_pdl.argtypes = [c_int]
        _lib.dlb_set_pdl.restype = None
        _lib.dlb_set_pdl(1 if os.environ.get("DLB_PDL", "1") == "1" else 0)     # programmatic dependent launch
    return _lib


def available() -> bool:
    """True when the native kernels can run (library present AND a CUDA device)."""
    return torch.cuda.is_available() and get() is not None


def require() -> ctypes.CDLL:
    lib = get()
    if lib is None:
        raise RuntimeError(f"native library {LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
    return lib


def stream_ptr(device=None) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"{what} failed with code {rc}")


def launch_count() -> int:
    lib = get()
    return int(lib.dlb_launch_count()) if lib is not None else 0


def ptr(t: Optional[torch.Tensor]) -> int:
    return 0 if t is None else t.data_ptr()
