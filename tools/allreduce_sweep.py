#!/usr/bin/env python
"""Weighted-allreduce bus-bandwidth sweep (BASELINE config #5): our fused one-shot / two-shot / NVLS kernels vs
the reference's path (scale kernel + NCCL all_reduce) from 1 KB to 1 GB, device-timed, max over ranks.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/allreduce_sweep.py

busbw = algbw * 2(n-1)/n (NCCL convention); roofline = measured 770 GB/s peer copy per direction (B200_PROFILING.md).
"""
import datetime
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dynamic_load_balance_distributeddnn_b200.parallel import SymmComm  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dev = f"cuda:{local}"
dist.init_process_group("nccl", device_id=torch.device(dev), timeout=datetime.timedelta(seconds=300))
max_bytes = int(os.environ.get("SWEEP_MAX_BYTES", str(1 << 30)))
n_max = max_bytes // 4
comm = SymmComm(dev, timeout_s=20.0)
comm.gate = False            # time the collective kernels alone (the one-warp bucket gate is a training-loop device)
gin, gout = comm.alloc_grad_buffers(n_max, torch.float32, dev)
gin.normal_()
w = torch.full((world,), 1.0 / world, device=dev)
ref_buf = torch.randn(n_max, device=dev)
flush = torch.empty(192 << 20, dtype=torch.uint8, device=dev)


def timed(fn, iters):
    fn(); fn()
    torch.cuda.synchronize(); dist.barrier()
    ts = []
    for _ in range(iters):
        flush.zero_()
        comm.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = torch.tensor([sorted(ts)[len(ts) // 2]], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


rows = []
size = 1024
algos = ["oneshot", "twoshot"] + (["nvls"] if comm.has_multicast else [])
while size <= max_bytes:
    n = size // 4
    n = max(32 * world, n // (32 * world) * (32 * world))
    iters = 20 if size <= (1 << 24) else 5
    res = {"bytes": n * 4}
    for algo in algos:
        if algo == "oneshot" and size > (1 << 26):
            continue
        comm.algo = algo
        ms = timed(lambda: comm.allreduce_buckets(gin, gout, [(0, n)], w), iters)
        res[algo + "_us"] = ms * 1e3
        res[algo + "_busbw"] = n * 4 / (ms * 1e-3) / 1e9 * 2 * (world - 1) / world
        if algo == "nvls":
            # training applies the DBS weight in the pack kernel, so the collective itself runs unweighted (no staging pass)
            ms = timed(lambda: comm.allreduce_buckets(gin, gout, [(0, n)], None), iters)
            res["nvls_noscale_us"] = ms * 1e3
            res["nvls_noscale_busbw"] = n * 4 / (ms * 1e-3) / 1e9 * 2 * (world - 1) / world
    scale = 1.0 / world

    def nccl_path():
        t = ref_buf[:n] * scale                      # the reference's separate `weighted * grad` kernel (dbs.py:295)
        dist.all_reduce(t)
    ms = timed(nccl_path, iters)
    res["nccl_scale_us"] = ms * 1e3
    res["nccl_scale_busbw"] = n * 4 / (ms * 1e-3) / 1e9 * 2 * (world - 1) / world
    rows.append(res)
    if rank == 0:
        print(json.dumps(res), flush=True)
    size *= 4
comm.check_errors()
if rank == 0:
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump({"world": world, "multicast": comm.has_multicast, "rows": rows},
              open(os.path.join(ROOT, "gpurun_out", f"allreduce_sweep_n{world}.json"), "w"))
comm.close()
dist.destroy_process_group()
