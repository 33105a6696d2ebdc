"""Process launcher / runtime init: one process per rank (per GPU).

Capabilities of reference ``dbs.py:511-544``: spawn ``world_size`` workers, rendezvous on
``127.0.0.1``, bind rank→device from ``-gpu``, per-rank logger, run the trainer.  Differences: runs
under ``torchrun`` too (uses RANK/WORLD_SIZE/LOCAL_RANK when present), propagates child exit codes and
tears the job down when one rank dies (the reference ignores them), and "already finished"
is a completion marker written at the end, not the mere existence of rank 0's log.
"""
from __future__ import annotations

import datetime
import os
import traceback

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from .config import DBSConfig
from .utils import done_marker, init_logger


def _backend_for(cfg: DBSConfig, from_env: bool = False) -> str:
    if cfg.debug or cfg.comm == "gloo":
        return "gloo"
    if from_env and isinstance(cfg.gpu, int):
        return "nccl"            # torchrun: one rank per GPU (LOCAL_RANK), regardless of the `-gpu` default
    gpus = cfg.gpu if isinstance(cfg.gpu, list) else [cfg.gpu] * cfg.world_size
    used = [gpus[r % len(gpus)] for r in range(cfg.world_size)]
    if len(set(used)) < len(used):
        return "gloo"            # several ranks share a GPU (reference `-gpu 0,0,0,1`): NCCL refuses duplicates
    return "nccl"


def worker(rank: int, world: int, cfg: DBSConfig, from_env: bool = False) -> None:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", str(cfg.master_port))
    backend = _backend_for(cfg, from_env)
    device = cfg.device_for_rank(rank)
    if from_env and not cfg.debug and isinstance(cfg.gpu, int) and world > 1:
        device = f"cuda:{int(os.environ.get('LOCAL_RANK', rank))}"
    if device.startswith("cuda"):
        torch.cuda.set_device(torch.device(device))
    else:                                   # CPU debug mode: do not oversubscribe the cores
        torch.set_num_threads(max(1, (os.cpu_count() or 1) // max(1, world)))
    kw = {}
    if backend == "nccl":
        kw["device_id"] = torch.device(device)
    dist.init_process_group(backend, rank=rank, world_size=world, timeout=datetime.timedelta(seconds=600), **kw)
    logger = init_logger(cfg, rank)
    try:
        from .engine import Trainer
        if backend == "gloo" and cfg.comm in ("auto", "symm", "nccl") and device.startswith("cuda"):
            cfg = cfg.replace(comm="gloo")
        trainer = Trainer(cfg, rank, world, device, logger)
        trainer.run()
        trainer.close()
        if rank == 0:
            with open(done_marker(cfg), "w") as f:
                f.write("done\n")
    except Exception:
        logger.error("Rank %d failed:\n%s", rank, traceback.format_exc())
        raise
    finally:
        try:
            dist.destroy_process_group()
        except Exception:
            pass


def _spawn_entry(rank: int, world: int, cfg: DBSConfig) -> None:
    worker(rank, world, cfg, from_env=False)


def launch(cfg: DBSConfig) -> int:
    """Returns a process exit code."""
    if os.path.isfile(done_marker(cfg)) and not cfg.force:
        print("\n===========================\nHad finished this experiments, skipping...\n===========================\n")
        return 0
    if "RANK" in os.environ and "WORLD_SIZE" in os.environ:            # torchrun / external launcher
        world = int(os.environ["WORLD_SIZE"])
        cfg = cfg.replace(world_size=world)
        worker(int(os.environ["RANK"]), world, cfg, from_env=True)
        return 0
    if not cfg.debug and isinstance(cfg.gpu, int) and cfg.world_size > 1 and torch.cuda.device_count() >= cfg.world_size \
            and os.environ.get("DLB_SPREAD_GPUS", "0") == "1":
        cfg = cfg.replace(gpu=list(range(cfg.world_size)))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(cfg.master_port)
    ctx = mp.get_context("spawn")
    procs = []
    for rank in range(cfg.world_size):
        p = ctx.Process(target=_spawn_entry, args=(rank, cfg.world_size, cfg), daemon=False)
        p.start()
        procs.append(p)
    code = 0
    alive = set(range(len(procs)))
    while alive:
        for i in list(alive):
            p = procs[i]
            p.join(timeout=0.2)
            if p.exitcode is not None:
                alive.discard(i)
                if p.exitcode != 0 and code == 0:
                    code = p.exitcode
                    for j in alive:                     # one rank died: do not leave peers hanging in a collective
                        procs[j].terminate()
    return code
