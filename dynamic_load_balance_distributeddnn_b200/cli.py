"""Command line: the reference's 13 flags with identical spellings, types and defaults
(reference ``parser.py:40-80``), plus opt-in extensions.

Differences, all deliberate: a bare ``-gpu 0`` is accepted (the reference
rejects it, ``parser.py:19-25``), invalid model/dataset names give a proper argparse error
instead of a ``TypeError`` from a mis-constructed ``ArgumentError``, and extra model names
(``resnet50`` …) are accepted on top of the reference six.
"""
from __future__ import annotations

import argparse
from typing import List, Optional, Sequence, Union

from .config import DATASETS, MODELS, DBSConfig


def str2bool(v) -> bool:
    """Same truthy/falsy vocabulary as the reference (``parser.py:8-16``)."""
    if isinstance(v, bool):
        return v
    s = str(v).lower()
    if s in ("yes", "true", "t", "y", "1"):
        return True
    if s in ("no", "false", "f", "n", "0"):
        return False
    raise argparse.ArgumentTypeError("Boolean value expected.")


def gpu_list(v) -> Union[int, List[int]]:
    """``0`` → int 0 (all ranks on that GPU); ``0,0,0,1`` → rank→GPU list."""
    if isinstance(v, int):
        return v
    try:
        if "," in v:
            return [int(g) for g in v.split(",") if g != ""]
        return int(v)
    except ValueError:
        raise argparse.ArgumentTypeError("Accepts GPU Number or GPU list")


def dataset_name(v: str) -> str:
    if v not in DATASETS:
        raise argparse.ArgumentTypeError("Invalid dataset (choose from %s)" % ", ".join(DATASETS))
    return v


def model_name(v: str) -> str:
    if v not in MODELS:
        raise argparse.ArgumentTypeError("Invalid model (choose from %s)" % ", ".join(MODELS))
    return v


def get_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(description="Dynamic Batchsize for Distributed DNN Training (B200-native)")
    # ---- reference flags -------------------------------------------------------------
    p.add_argument("-d", "--debug", type=str2bool, default=True,
                   help="Debug mode: run on CPU over gloo. Default True.")
    p.add_argument("-ws", "--world_size", type=int, default=4, help="Number of workers. Default 4.")
    p.add_argument("-b", "--batch_size", type=int, default=64,
                   help="GLOBAL batch size of the cluster (re-split across workers by DBS). Default 64.")
    p.add_argument("-lr", "--learning_rate", type=float, default=0.01, help="SGD learning rate. Default 0.01.")
    p.add_argument("-e", "--epoch_size", type=int, default=10, help="Number of epochs. Default 10.")
    p.add_argument("-ds", "--dataset", type=dataset_name, default="wikitext2",
                   help="cifar10 | cifar100 | mnist (FashionMNIST) | wikitext2. Default wikitext2.")
    p.add_argument("-dbs", "--dynamic_batch_size", type=str2bool, default=True,
                   help="Enable the dynamic-batch-size rebalancer. Default True.")
    p.add_argument("-gpu", "--gpu", type=gpu_list, default=0,
                   help="GPU id, or comma list mapping rank→GPU, e.g. 0,0,0,1. Ignored in debug mode.")
    p.add_argument("-m", "--model", type=model_name, default="transformer",
                   help="mnistnet | resnet (=ResNet-101) | densenet (=DenseNet-121) | googlenet | regnet "
                        "(=RegNetY-400MF) | transformer, plus explicit variants (resnet50, densenet169, …).")
    p.add_argument("-ft", "--fault_tolerance", type=str2bool, default=False,
                   help="Inject random stragglers (per-epoch Bernoulli, multi-epoch slow phase). Default False.")
    p.add_argument("-ftc", "--fault_tolerance_chance", type=float, default=0.1,
                   help="Per-epoch probability that a worker becomes a straggler. Default 0.1.")
    p.add_argument("-ocp", "--one_cycle_policy", type=str2bool, default=False,
                   help="Enable the one-cycle learning-rate policy.")
    p.add_argument("-de", "--disable_enhancements", type=str2bool, default=False,
                   help="Ablation: uniform 1/world_size gradient weights and no LR policy.")
    # ---- extensions ------------------------------------------------------------------
    x = p.add_argument_group("extensions (defaults reproduce the reference behaviour)")
    x.add_argument("--seed", type=int, default=1234)
    x.add_argument("--comm", choices=("auto", "gloo", "nccl", "symm"), default="auto",
                   help="gradient transport: symm = fused sm_100a P2P/NVLS kernels (default on GPU), "
                        "nccl = A/B baseline, gloo = CPU debug")
    x.add_argument("--dtype", choices=("auto", "fp32", "tf32", "bf16"), default="auto",
                   help="bf16 (default on GPU): bf16 compute, fp32 master weights/accumulation; tf32 / fp32: fp32 storage, TF32 "
                        "tensor-core math (the reference's precision class)")
    x.add_argument("--synthetic", type=str2bool, default=None,
                   help="force synthetic data of the dataset's shape (auto when files are absent)")
    x.add_argument("--train_samples", type=int, default=0)
    x.add_argument("--test_samples", type=int, default=0)
    x.add_argument("--data_root", default="./data")
    x.add_argument("--corpus_root", default="")
    x.add_argument("--log_dir", default="./logs")
    x.add_argument("--stats_dir", default="./statis")
    x.add_argument("--rounding", choices=("largest_remainder", "reference"), default="largest_remainder")
    x.add_argument("--min_local_batch", type=int, default=1)
    x.add_argument("--batch_quantum", type=int, default=1)
    x.add_argument("--rebalance_every", type=int, default=0,
                   help="rebalance every N steps instead of once per epoch (0 = per epoch)")
    x.add_argument("--time_ema", type=float, default=0.0)
    x.add_argument("--dbs_model", choices=("auto", "proportional", "affine"), default="auto",
                   help="proportional = the reference's rule; affine = fit t_r(b) = alpha + beta*b per rank and equalise "
                        "predicted step times (for latency-bound steps where time is not proportional to the batch)")
    x.add_argument("--lr_policy", choices=("one_cycle", "legacy"), default="one_cycle")
    x.add_argument("--clip_grad_norm", type=float, default=-1.0)
    x.add_argument("--clip_mode", choices=("local", "global"), default="local")
    x.add_argument("--throttle_rank", type=int, default=-1)
    x.add_argument("--throttle_ms", type=float, default=0.0)
    x.add_argument("--throttle_mode", choices=("sleep", "burn"), default="sleep")
    x.add_argument("--cuda_graphs", type=str2bool, default=True)
    x.add_argument("--bucket_mb", type=float, default=8.0)
    x.add_argument("--overlap_comm", type=str2bool, default=True, help="overlap bucket allreduces with backward")
    x.add_argument("--wire_dtype", choices=("fp32", "bf16"), default="fp32")
    x.add_argument("--allreduce_algo", choices=("auto", "oneshot", "twoshot", "nvls"), default="auto")
    x.add_argument("--comm_timeout_s", type=float, default=20.0, help="device-side watchdog of the fused collectives")
    x.add_argument("--max_steps_per_epoch", type=int, default=0)
    x.add_argument("--validate", type=str2bool, default=True)
    x.add_argument("--bptt", type=int, default=35)
    x.add_argument("--checkpoint_dir", default="")
    x.add_argument("--resume", type=str2bool, default=False)
    x.add_argument("--force", type=str2bool, default=False)
    x.add_argument("--master_port", type=int, default=29500)
    x.add_argument("--profile", type=str2bool, default=False)
    return p


def config_from_args(argv: Optional[Sequence[str]] = None) -> DBSConfig:
    ns = get_parser().parse_args(argv)
    return DBSConfig(**vars(ns))


def main(argv: Optional[Sequence[str]] = None) -> int:
    from .launch import launch
    cfg = config_from_args(argv)
    return launch(cfg)
