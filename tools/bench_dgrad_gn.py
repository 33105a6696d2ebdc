#!/usr/bin/env python
"""Micro-benchmark of the GN1 backward of a DenseNet bottleneck on its real shapes: the default three-kernel chain
(tcgen05 dgrad GEMM -> nc_reduce2<coef mask> -> gn_bwd_apply<coef mask>, 7 x |x| bytes) against the experimental fused
two-pass kernel of csrc/dgrad_gn.cu (4 x |x| bytes).  CUDA-event timed, L2 flushed between iterations.

    python tools/bench_dgrad_gn.py [quick]
"""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dynamic_load_balance_distributeddnn_b200.ops import _native as nat, gemm_tc  # noqa: E402
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
lib = nat.require()
groups, cm = 32, 128
# (samples, H*W, Cin of the layer, total channels of the block buffer): first / middle / last layer of each dense block at B = 512
shapes = [(512, 1024, 64, 256), (512, 1024, 224, 256), (512, 256, 128, 512), (512, 256, 480, 512), (512, 64, 256, 1024),
          (512, 64, 992, 1024), (64, 1024, 224, 256), (64, 256, 480, 512)]
if quick:
    shapes = [(512, 256, 480, 512)]
for (ns, hw, cl, ct) in shapes:
    m = ns * hw
    dy = torch.randn(m, cm, device="cuda").bfloat16()
    w = (torch.randn(cm, cl, device="cuda") / cm ** 0.5).bfloat16()
    buf = torch.randn(m, ct, device="cuda").bfloat16(); dbuf = torch.randn(m, ct, device="cuda").bfloat16()
    off = ct - cl
    esz = 2
    xs, dxs = buf.data_ptr() + off * esz, dbuf.data_ptr() + off * esz
    kp = (cl + 63) // 64 * 64
    ca = torch.zeros(ns, kp, device="cuda"); cb = torch.zeros(ns, kp, device="cuda")
    ca[:, :cl] = torch.rand(ns, cl, device="cuda") + 0.5; cb[:, :cl] = torch.randn(ns, cl, device="cuda") * 0.3
    gamma = torch.rand(cl, device="cuda") + 0.5
    mean = torch.randn(ns * groups, device="cuda") * 0.1; rstd = torch.rand(ns * groups, device="cuda") + 0.5
    t1 = torch.zeros(ns * cl * 2, device="cuda"); dg = torch.zeros(cl, device="cuda"); db = torch.zeros(cl, device="cuda")
    dxhat = torch.empty(m, cl, device="cuda", dtype=torch.bfloat16)
    k23 = torch.empty(2, ns, kp, device="cuda")
    st = nat.stream_ptr(buf.device)

    def chain():
        gemm_tc.gemm_bmn_raw(dy.data_ptr(), cm, w.data_ptr(), cl, dxhat.data_ptr(), cl, m, cl, cm, buf.device)
        nat.check(lib.dlb_nc_reduce2_bwd_coef(1, xs, ct, dxhat.data_ptr(), cl, t1.data_ptr(), 0, mean.data_ptr(), rstd.data_ptr(),
                                              dg.data_ptr(), db.data_ptr(), ca.data_ptr(), cb.data_ptr(), kp, ns, hw, cl, groups, st), "red")
        nat.check(lib.dlb_gn_bwd_apply_coef(1, xs, ct, dxhat.data_ptr(), cl, dxs, ct, gamma.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                            t1.data_ptr(), 0, ca.data_ptr(), cb.data_ptr(), kp, ns, hw, cl, groups, 1, st), "app")

    def fused():
        t1.zero_()
        gemm_tc.dgrad_gn_raw(1, dy.data_ptr(), cm, w.data_ptr(), cl, xs, ct, 0, 0, m, cl, cm, hw, ca, cb, None, None, t1.data_ptr(), 2 * cl, buf.device)
        gemm_tc.gn_bwd_coeff_raw(t1.data_ptr(), 2 * cl, gamma, mean, rstd, k23[0], k23[1], dg.data_ptr(), db.data_ptr(), ns, cl, groups, hw, buf.device)
        gemm_tc.dgrad_gn_raw(2, dy.data_ptr(), cm, w.data_ptr(), cl, xs, ct, dxs, ct, m, cl, cm, hw, ca, cb, k23[0], k23[1], 0, 0, buf.device)

    def pass1():
        gemm_tc.dgrad_gn_raw(1, dy.data_ptr(), cm, w.data_ptr(), cl, xs, ct, 0, 0, m, cl, cm, hw, ca, cb, None, None, t1.data_ptr(), 2 * cl, buf.device)

    def pass2():
        gemm_tc.dgrad_gn_raw(2, dy.data_ptr(), cm, w.data_ptr(), cl, xs, ct, dxs, ct, m, cl, cm, hw, ca, cb, k23[0], k23[1], 0, 0, buf.device)

    tc = timeit(chain)
    line = f"N={ns} HW={hw} Cin={cl}: chain {tc*1e3:7.1f} us ({7*m*cl*2/tc/1e6:5.0f} GB/s over 7|x|)"
    if gemm_tc.dgrad_gn_available() and hw % 32 == 0:
        tf, tp1, tp2 = timeit(fused), timeit(pass1), timeit(pass2)
        x1 = (m * cl + m * cm) * 2; x2 = (3 * m * cl + m * cm) * 2
        line += (f" | fused {tf*1e3:7.1f} us = pass1 {tp1*1e3:6.1f} ({x1/tp1/1e6:5.0f} GB/s, {100*x1/tp1/1e6/PEAK:4.1f}%) + pass2 {tp2*1e3:6.1f} "
                 f"({x2/tp2/1e6:5.0f} GB/s, {100*x2/tp2/1e6/PEAK:4.1f}% of measured copy peak)  speedup {tc/tf:4.2f}x")
    print(line, flush=True)
