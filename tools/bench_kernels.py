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


def gnbwd():
    """fused GroupNorm backward (one launch: reduce + per-sample barrier + apply): register flavour vs bulk-copy flavour at
    several tile budgets, DenseNet-121 shapes (x / dX = channel slices of the block buffer, accumulate) and the bottleneck
    shape (C = 128, fresh output).  Traffic model: x + dy read once, dX read + written (4 passes; 3 without accumulation)."""
    lib = nat.require()
    st = nat.stream_ptr()
    out = []
    shapes = [(512, 1024, 128, 128, 0), (512, 1024, 224, 256, 1), (512, 256, 128, 128, 0), (512, 256, 480, 512, 1), (512, 64, 128, 128, 0),
              (512, 64, 992, 1024, 1), (64, 1024, 224, 256, 1), (64, 256, 480, 512, 1), (64, 64, 992, 1024, 1), (64, 64, 128, 128, 0)]
    for dname, dtype in (("tf32", torch.float32), ("bf16", torch.bfloat16)):
        for (n, hw, c, ct, acc) in shapes:
            xb = torch.randn(n, hw, ct, device="cuda").to(dtype); db_ = torch.randn(n, hw, ct, device="cuda").to(dtype)
            x, dx = xb[..., ct - c:], db_[..., ct - c:]
            dy = torch.randn(n, hw, c, device="cuda").to(dtype)
            gamma = torch.ones(c, device="cuda"); mean = torch.zeros(n * 32, device="cuda"); rstd = torch.ones(n * 32, device="cuda")
            kp = (c + 63) // 64 * 64
            ca = torch.ones(n, kp, device="cuda"); cb = torch.zeros(n, kp, device="cuda")
            table = torch.zeros(n, 2 * c, device="cuda"); dg = torch.zeros(c, device="cuda"); dbt = torch.zeros(c, device="cuda")
            done = torch.zeros(n, dtype=torch.int32, device="cuda")
            nbytes = n * hw * c * x.element_size()
            traffic = (4 if acc else 3) * nbytes

            def run():
                done.zero_()
                rc = lib.dlb_gn_bwd_fused(nat.dtype_code(dtype), x.data_ptr(), ct, dy.data_ptr(), c, dx.data_ptr(), ct, gamma.data_ptr(),
                                          mean.data_ptr(), rstd.data_ptr(), table.data_ptr(), 0, dg.data_ptr(), dbt.data_ptr(),
                                          ca.data_ptr(), cb.data_ptr(), kp, done.data_ptr(), n, hw, c, 32, acc, st)
                assert rc in (0, 1), rc
                return rc
            line = f"{dname} N={n:4d} HW={hw:5d} C={c:4d}/{ct:4d} acc={acc} ({traffic / 1e6:7.1f} MB):"
            for tag, on, kb in (("reg", 0, 0), ("bulk32", 1, 32), ("bulk48", 1, 48), ("bulk64", 1, 64), ("bulk96", 1, 96)):
                lib.dlb_norm_bulk(on, kb)
                if run() == 1:
                    line += f"  {tag} n/a"
                    continue
                ms = timeit(run)
                gbs = traffic / ms / 1e6
                out.append((dname, n, hw, c, ct, acc, tag, ms * 1e3, gbs / PEAK))
                line += f"  {tag} {ms * 1e3:6.1f}us {100 * gbs / PEAK:4.0f}%"
            lib.dlb_norm_bulk(0, 64)
            print(line, flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "gn_bwd_flavours.json"), "w"))


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "gnbwd":
        return gnbwd()
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
