"""One forward+backward of a DenseNet-121 block-1-shaped dense stage in eager mode, for `ncu --set full` captures of the
kernels that make up the step (GN-prologue GEMM, halo 3x3 conv fwd/dgrad, 9-tap wgrad, split-K wgrad, fused dgrad+GN
backward, fused GN backward, GN apply).  The first pass warms up; the second runs between cudaProfilerStart/Stop:

    ncu --set full --clock-control none --import-source on --profile-from-start off -c 40 -o out \
        python tools/ncu_targets.py --dtype tf32 --batch 256 --layers 3

Reference shapes: Net/Densenet.py:9-30 (bottleneck 4*growth = 128 mid channels, growth 32, 32x32 maps in block 1).
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dtype", default="tf32", choices=("tf32", "bf16"))
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--layers", type=int, default=3)
    ap.add_argument("--hw", type=int, default=32)
    ap.add_argument("--c0", type=int, default=64)
    a = ap.parse_args()
    from dynamic_load_balance_distributeddnn_b200.models import densenet
    dt = torch.bfloat16 if a.dtype == "bf16" else torch.float32
    torch.manual_seed(0)
    net = densenet.DenseNet([a.layers], growth_rate=32, num_classes=10)
    stage = net.dense1.cuda()
    if a.c0 != 64:
        raise SystemExit("c0 is fixed by the model's stem (64)")
    for m in stage.modules():
        if isinstance(m, torch.nn.Conv2d):
            m.weight.data = m.weight.data.to(dt)
    x0 = torch.randn(a.batch, 64, a.hw, a.hw, device="cuda").contiguous(memory_format=torch.channels_last).to(dt)
    for it in range(2):
        if it == 1:
            torch.cuda.synchronize()
            torch.cuda.profiler.start()
        for p in stage.parameters():
            p.grad = None
        x = x0.clone().requires_grad_(True)
        y = stage(x)
        y.backward(torch.ones_like(y))
        torch.cuda.synchronize()
        if it == 1:
            torch.cuda.profiler.stop()
    print("ok", tuple(y.shape), a.dtype)


if __name__ == "__main__":
    main()
