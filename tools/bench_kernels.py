#!/usr/bin/env python
"""Micro-benchmark of the hand-written memory-bound kernels on DenseNet-121 shapes (B=512) against the
measured copy bandwidth (MEASURED_PEAKS.json hbm_gbs).  CUDA-event timed, L2 flushed between iterations."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dynamic_load_balance_distributeddnn_b200 import ops  # noqa: E402
from dynamic_load_balance_distributeddnn_b200.ops import _native as nat  # noqa: E402

PEAK = 6462.4
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass

flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, iters=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


def main():
    lib = nat.require()
    st = nat.stream_ptr()
    rows = []
    quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
    sweep = len(sys.argv) > 1 and sys.argv[1] == "sweep"
    import ctypes
    if hasattr(lib, "dlb_norm_tune"):
        lib.dlb_norm_tune.argtypes = [ctypes.c_int] * 3
        lib.dlb_norm_tune_kb.argtypes = [ctypes.c_int]
    shapes = [(512, 1024, 64), (512, 1024, 128), (512, 1024, 256), (512, 256, 128), (512, 256, 512), (512, 64, 1024), (512, 16, 1024),
              (64, 1024, 128), (64, 256, 512)]
    if quick:
        shapes = [(512, 1024, 128), (512, 64, 1024)]
    if sweep:
        shapes = [(512, 1024, 128), (512, 256, 256), (512, 64, 640), (64, 1024, 128), (64, 64, 640)]
    configs = [(0, 0)] if not sweep else [(u, kb) for u in (1, 2, 4) for kb in (48, 96, 192)]
    for (unr, kb) in configs:
      if sweep:
        lib.dlb_norm_tune(unr, unr, unr); lib.dlb_norm_tune_kb(kb)
        print(f"---- rows in flight per thread = {unr}, min KB per block = {kb}", flush=True)
      for (n, hw, c) in shapes:
          h = int(hw ** 0.5)
          x = torch.randn(n, c, h, h, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last)
          y = torch.empty_like(x); dy = torch.randn_like(x); dx = torch.empty_like(x)
          w = torch.ones(c, device="cuda"); b = torch.zeros(c, device="cuda")
          mean = torch.zeros(n * 32, device="cuda"); rstd = torch.ones(n * 32, device="cuda")
          table = torch.empty(n * c * 2, device="cuda"); dg = torch.empty(c, device="cuda"); db = torch.empty(c, device="cuda")
          nbytes = x.numel() * 2
          def f_reduce(): lib.dlb_nc_reduce2(0, 1, x.data_ptr(), c, 0, 0, 0, 0, table.data_ptr(), 0, n, hw, c, st)
          def f_reduce_b(): lib.dlb_nc_reduce2(1, 1, x.data_ptr(), c, dy.data_ptr(), c, y.data_ptr(), c, table.data_ptr(), 0, n, hw, c, st)
          def f_apply(): lib.dlb_gn_fwd_apply(1, x.data_ptr(), c, 0, 0, y.data_ptr(), c, w.data_ptr(), b.data_ptr(), mean.data_ptr(), rstd.data_ptr(), n, hw, c, 32, 1, st)
          def f_bapply(): lib.dlb_gn_bwd_apply(1, x.data_ptr(), c, dy.data_ptr(), c, y.data_ptr(), c, dx.data_ptr(), c, 0, 0, w.data_ptr(), mean.data_ptr(), rstd.data_ptr(), table.data_ptr(), 0, n, hw, c, 32, 1, 0, st)
          def f_pgrad(): lib.dlb_gn_param_grad(table.data_ptr(), 0, mean.data_ptr(), rstd.data_ptr(), dg.data_ptr(), db.data_ptr(), n, c, 32, st)
          def f_copy(): y.copy_(x)
          for name, fn, traffic in (("copy(torch)", f_copy, 2 * nbytes), ("nc_reduce2 fwd", f_reduce, nbytes), ("nc_reduce2 bwd", f_reduce_b, 3 * nbytes),
                                    ("gn_fwd_apply", f_apply, 2 * nbytes), ("gn_bwd_apply", f_bapply, 4 * nbytes),
                                    ("gn_param_grad", f_pgrad, n * c * 8)):
              ms = timeit(fn)
              gbs = traffic / ms / 1e6
              rows.append((n, hw, c, name, ms * 1e3, gbs, gbs / PEAK, unr, kb))
              print(f"N={n:4d} HW={hw:5d} C={c:5d} {name:16s} {ms * 1e3:9.1f} us {gbs:8.0f} GB/s  {100 * gbs / PEAK:5.1f}% of measured copy peak", flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "kernel_bench.json"), "w"))


if __name__ == "__main__":
    main()
