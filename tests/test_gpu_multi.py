"""Multi-GPU tests of the fused collectives (2+ GPUs; ``gpurun --gpus 2``)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, datetime, torch, torch.distributed as dist
sys.path.insert(0, %(root)r)
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{rank}"), timeout=datetime.timedelta(seconds=120))
from dynamic_load_balance_distributeddnn_b200.parallel import SymmComm
c = SymmComm(f"cuda:{rank}", timeout_s=20.0)
n = 1 << 20
for wire in (torch.float32, torch.bfloat16):
    gin, gout = c.alloc_grad_buffers(n, wire, f"cuda:{rank}")
    torch.manual_seed(100 + rank)
    local = torch.randn(n, device="cuda")
    w = torch.tensor([(r + 1.0) for r in range(world)], device="cuda"); w /= w.sum()
    # reference result with NCCL
    ref = (local.to(wire).float() * w[rank]).clone()
    dist.all_reduce(ref)
    algos = ["oneshot", "twoshot"] + (["nvls"] if c.has_multicast else [])
    for algo in algos:
        c.algo = algo
        gin.copy_(local.to(wire)); gout.zero_()
        torch.cuda.synchronize(); dist.barrier()
        c.allreduce_buckets(gin, gout, [(0, n // 2), (n // 2, n // 2)], w)      # weights applied in-kernel
        torch.cuda.synchronize()
        c.check_errors()
        tol = 1e-5 if wire == torch.float32 else 2e-2
        err = (gout.float() - ref).abs().max().item()
        assert err < tol * max(1.0, ref.abs().max().item()), (algo, wire, err)
        # replicas must be bit-identical
        chk = gout.float().double().sum().reshape(1).clone()
        lst = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(lst, chk)
        assert all(torch.equal(lst[0], x) for x in lst), (algo, wire)
        if rank == 0: print("ok", algo, wire, "err", err, flush=True)
        # unweighted flavour (what training uses: the DBS weight is applied by the pack kernel)
        gin.copy_((local * w[rank]).to(wire)); gout.zero_()
        torch.cuda.synchronize(); dist.barrier()
        c.allreduce_buckets(gin, gout, [(0, n)], None)
        torch.cuda.synchronize(); c.check_errors()
        ref2 = (local * w[rank]).to(wire).float().clone(); dist.all_reduce(ref2)
        err2 = (gout.float() - ref2).abs().max().item()
        assert err2 < tol * max(1.0, ref2.abs().max().item()), (algo, wire, "unweighted", err2)
# ---- optimizer step fused behind the collective: allreduce + momentum SGD + bf16 shadow + clearing of the input, ONE launch
n2 = (1 << 18) + 32 * world
gin, gout = c.alloc_grad_buffers(n2, torch.float32, f"cuda:{rank}")
for algo in ["oneshot", "twoshot"] + (["nvls"] if c.has_multicast else []):
    c.algo = algo
    torch.manual_seed(7)
    master = torch.randn(n2, device="cuda"); mom = torch.randn(n2, device="cuda") * 0.1
    shadow = torch.zeros(n2, device="cuda", dtype=torch.bfloat16)
    lr = torch.full((1,), 0.05, device="cuda")
    torch.manual_seed(200 + rank)
    local = torch.randn(n2, device="cuda")
    ref_g = local.clone(); dist.all_reduce(ref_g)
    ref_m = 0.9 * mom + ref_g
    ref_p = master - 0.05 * ref_m
    gin.copy_(local); gout.zero_()
    torch.cuda.synchronize(); dist.barrier()
    c.allreduce_buckets_sgd(gin, gout, [(0, n2 // 2), (n2 // 2, n2 - n2 // 2)],
                            {"master": master.data_ptr(), "mom": mom.data_ptr(), "shadow": shadow.data_ptr(), "lr": lr.data_ptr(),
                             "momentum": 0.9, "weight_decay": 0.0, "zero_in": gin.data_ptr()})
    torch.cuda.synchronize(); c.check_errors()
    assert (master - ref_p).abs().max().item() < 1e-4, (algo, (master - ref_p).abs().max().item())
    assert (mom - ref_m).abs().max().item() < 1e-4, algo
    assert (shadow.float() - ref_p).abs().max().item() < 2e-2 * ref_p.abs().max().item(), algo
    assert gin.abs().max().item() == 0.0, (algo, "gradient buffer not cleared")
    chk = master.double().sum().reshape(1).clone()
    lst = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(lst, chk)
    assert all(torch.equal(lst[0], x) for x in lst), (algo, "replicas diverged")
    if rank == 0: print("ok fused-sgd", algo, flush=True)
t = c.gather_times(10.0 + rank)
assert t == [10.0 + r for r in range(world)], t
t = c.gather_times(20.0 + rank)
assert t == [20.0 + r for r in range(world)], t
c.barrier(); torch.cuda.synchronize()
if rank == 0: print("backend", c.alloc.buffers["grad_in"].backend, "multicast", c.has_multicast, flush=True)
c.close()
dist.destroy_process_group()
'''


def _torchrun(script_path, nproc, port, timeout=300):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), script_path]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)


def test_weighted_allreduce_all_algos(tmp_path):
    n = min(torch.cuda.device_count(), 8)
    p = tmp_path / "w.py"
    p.write_text(WORKER % {"root": ROOT})
    r = _torchrun(str(p), n, 29701)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "ok twoshot" in r.stdout and "ok fused-sgd twoshot" in r.stdout


def test_end_to_end_two_ranks_rebalance(tmp_path):
    n = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29702", os.path.join(ROOT, "dbs.py"), "-d", "false", "-ws", str(n), "-b", "128", "-m", "resnet18",
           "-ds", "cifar10", "-e", "3", "--synthetic", "true", "--train_samples", "2560", "--test_samples", "256",
           "--throttle_rank", "1", "--throttle_ms", "15", "--log_dir", str(tmp_path / "logs"), "--stats_dir", str(tmp_path / "statis")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=str(tmp_path), env=dict(os.environ, PYTHONPATH=ROOT))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    from dynamic_load_balance_distributeddnn_b200.utils import load_stats
    import glob
    st = load_stats(glob.glob(str(tmp_path / "statis" / "*.npy"))[0])
    lb = st["local_batches"]
    assert lb[0] == [64, 64] and lb[-1][1] < 64 and sum(lb[-1]) == 128
