#!/usr/bin/env python
"""tcgen05 GEMM micro-benchmark on the DenseNet 1x1-conv shapes vs cuBLAS (torch.matmul), CUDA-event timed,
L2 flushed; reports achieved DRAM bandwidth against the measured copy peak (these GEMMs are memory-bound)."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dynamic_load_balance_distributeddnn_b200.ops import gemm_tc  # noqa: E402
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.isfile(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6462.4
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, iters=5):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)


quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
shapes = [(524288, 128, 64), (524288, 128, 256), (131072, 128, 512), (32768, 128, 1024), (8192, 128, 1024), (524288, 256, 128)]
if quick:
    shapes = [(524288, 128, 256)]
for (m, n, k) in shapes:
    a = torch.randn(m, k, device="cuda").bfloat16(); b = torch.randn(n, k, device="cuda").bfloat16()
    d = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
    t1 = timeit(lambda: gemm_tc.gemm(a, b, out=d))
    t2 = timeit(lambda: torch.matmul(a, b.t(), out=d))
    byts = (m * k + n * k + m * n) * 2
    print(f"M={m} N={n} K={k}: tcgen05 {t1*1e3:8.1f} us ({byts/t1/1e6:6.0f} GB/s, {100*byts/t1/1e6/PEAK:5.1f}% of measured copy peak, "
          f"{2*m*n*k/t1/1e9:7.1f} TFLOP/s)   cuBLAS {t2*1e3:8.1f} us ({byts/t2/1e6:6.0f} GB/s)", flush=True)

# ---- fused variants on DenseNet shapes: GN+ReLU prologue (+ stats epilogue), and the MN-major wgrad ----
print("-- fused variants")
fshapes = [(512, 1024, 128, 256), (512, 256, 128, 512), (512, 64, 128, 1024)] if not quick else [(512, 1024, 128, 256)]
for (ns, hw, n, k) in fshapes:
    m = ns * hw
    a = torch.randn(m, k, device="cuda").bfloat16(); b = torch.randn(n, k, device="cuda").bfloat16()
    d = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
    pa = torch.rand(ns, k, device="cuda") + 0.5; pb = torch.randn(ns, k, device="cuda") * 0.1
    table = torch.zeros(ns, n, 2, device="cuda")
    dy = torch.randn(m, n, device="cuda").bfloat16()
    byts = (m * k + n * k + m * n) * 2
    t0 = timeit(lambda: gemm_tc.gemm(a, b, out=d))
    t1 = timeit(lambda: gemm_tc.gemm(a, b, out=d, pro_a=pa, pro_b=pb, rows_per_sample=hw))
    t2 = timeit(lambda: gemm_tc.gemm(a, b, out=d, pro_a=pa, pro_b=pb, rows_per_sample=hw, stats=table, stats_ns=2 * n))
    t3 = timeit(lambda: gemm_tc.gemm(a, b, out=d, rows_per_sample=hw, stats=table, stats_ns=2 * n))
    dw = torch.zeros(n, k, device="cuda")
    t4 = timeit(lambda: gemm_tc.wgrad_raw(dy.data_ptr(), n, a.data_ptr(), k, dw, m, n, k, a.device))
    t5 = timeit(lambda: gemm_tc.wgrad_raw(dy.data_ptr(), n, a.data_ptr(), k, dw, m, n, k, a.device, pa, pb, hw))
    t6 = timeit(lambda: torch.matmul(dy.t(), a))
    print(f"N={ns} HW={hw} Cout={n} Cin={k}: plain {t0*1e3:6.1f}  +pro {t1*1e3:6.1f}  +pro+stats {t2*1e3:6.1f}  +stats {t3*1e3:6.1f} us "
          f"(fwd bytes {byts/1e6:.0f} MB -> {byts/t2/1e6:5.0f} GB/s fused) | wgrad {t4*1e3:6.1f}  +pro {t5*1e3:6.1f}  cuBLAS {t6*1e3:6.1f} us", flush=True)

# ---- tcgen05 implicit-GEMM 3x3 (DenseNet 128 -> 32 bottleneck conv) vs the vendor library ----
print("-- conv3x3 128->32")
import torch.nn.functional as F
for (n, hw) in ([(512, 32), (512, 16), (512, 8), (512, 4), (64, 32), (64, 16)] if not quick else [(512, 32)]):
    x = torch.randn(n, 128, hw, hw, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last)
    w = (torch.randn(32, 128, 3, 3, device="cuda") / 34).bfloat16().contiguous(memory_format=torch.channels_last)
    y = torch.empty(n, 32, hw, hw, device="cuda", dtype=torch.bfloat16, memory_format=torch.channels_last)
    dy = torch.randn_like(y); dx = torch.empty_like(x)
    t1 = timeit(lambda: gemm_tc.conv3x3_raw(False, x.data_ptr(), 128, w.data_ptr(), y.data_ptr(), 32, n, hw, hw, 128, 32, x.device))
    t2 = timeit(lambda: F.conv2d(x, w, padding=1))
    t3 = timeit(lambda: gemm_tc.conv3x3_raw(True, dy.data_ptr(), 32, w.data_ptr(), dx.data_ptr(), 128, n, hw, hw, 128, 32, x.device))
    t4 = timeit(lambda: torch.ops.aten.convolution_backward(dy, x, w, None, [1, 1], [1, 1], [1, 1], False, [0, 0], 1, [True, False, False]))
    print(f"N={n} {hw}x{hw}: fwd tcgen05 {t1*1e3:6.1f} us vs cuDNN {t2*1e3:6.1f} | dgrad tcgen05 {t3*1e3:6.1f} us vs cuDNN {t4*1e3:6.1f}", flush=True)
