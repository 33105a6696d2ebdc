"""Experiment configuration (one immutable dataclass instead of module globals).

The reference parses ``sys.argv`` at import time and unpacks the result into module
globals that every function reaches through ``global`` (reference ``dbs.py:22-44``,
``dbs.py:95,194,314``).  Here the CLI produces a :class:`DBSConfig` that is
passed explicitly through launcher → trainer → balancer, so spawned workers never
re-parse the command line and tests can build configs directly.
"""
from __future__ import annotations

import dataclasses
from dataclasses import dataclass
from typing import List, Optional, Union

MODELS = ("mnistnet", "resnet", "resnet18", "resnet34", "resnet50", "resnet101", "resnet152",
          "densenet", "densenet121", "densenet169", "densenet201", "densenet161",
          "googlenet", "regnet", "regnetx200", "regnetx400", "regnety400", "transformer")
#: the six model names the reference CLI accepts (reference ``parser.py:4``)
REFERENCE_MODELS = ("mnistnet", "resnet", "densenet", "googlenet", "regnet", "transformer")
DATASETS = ("cifar10", "cifar100", "mnist", "wikitext2")


@dataclass
class DBSConfig:
    # --- the 13 reference flags (reference ``parser.py:40-80``) ---
    debug: bool = True
    world_size: int = 4
    batch_size: int = 64                 # GLOBAL batch across the cluster
    learning_rate: float = 0.01
    epoch_size: int = 10
    dataset: str = "wikitext2"
    dynamic_batch_size: bool = True
    gpu: Union[int, List[int]] = 0
    model: str = "transformer"
    fault_tolerance: bool = False
    fault_tolerance_chance: float = 0.1
    one_cycle_policy: bool = False
    disable_enhancements: bool = False
    # --- extensions (all default to reference behaviour) ---
    seed: int = 1234
    momentum: float = 0.9
    comm: str = "auto"                   # auto | gloo | nccl | symm
    dtype: str = "auto"                  # auto (bf16 on cuda, fp32 on cpu) | fp32 | bf16
    synthetic: Optional[bool] = None     # None = auto (synthetic when the dataset is absent)
    train_samples: int = 0               # synthetic dataset size override (0 = dataset default)
    test_samples: int = 0
    data_root: str = "./data"
    corpus_root: str = ""                # wikitext-2 directory ("" = search, else synthetic)
    log_dir: str = "./logs"
    stats_dir: str = "./statis"
    rounding: str = "largest_remainder"  # largest_remainder | reference
    min_local_batch: int = 1
    batch_quantum: int = 1
    rebalance_every: int = 0             # 0 = once per epoch (reference); N>0 = every N steps
    time_ema: float = 0.0                # EMA on per-rank compute time (0 = off = reference)
    dbs_model: str = "auto"              # proportional (reference get_size) | affine (t = alpha + beta*b, latency-aware) | auto
                                         # (affine on CUDA devices, where a step has a large fixed cost; the reference rule on CPU)
    lr_policy: str = "one_cycle"         # one_cycle | legacy (the reference's live decay-only curve)
    clip_grad_norm: float = -1.0         # <0: model default (0.25 for transformer, none for CNNs)
    clip_mode: str = "local"             # local (reference, pre-allreduce) | global (post-reduce)
    throttle_rank: int = -1              # deterministic straggler: which rank
    throttle_ms: float = 0.0             # ... extra ms per step
    throttle_mode: str = "sleep"         # sleep | burn (device burner kernel)
    cuda_graphs: bool = True
    bucket_mb: float = 8.0
    overlap_comm: bool = True            # fire bucket collectives from autograd hooks on a side stream
    wire_dtype: str = "fp32"             # fp32 | bf16 wire format of the gradient allreduce
    allreduce_algo: str = "auto"         # auto | oneshot | twoshot | nvls
    comm_timeout_s: float = 20.0         # device-side watchdog of the fused collectives (a dead peer raises instead of hanging)
    max_steps_per_epoch: int = 0         # cap (0 = full epoch)
    validate: bool = True
    bptt: int = 35
    checkpoint_dir: str = ""
    resume: bool = False
    force: bool = False                  # ignore the "already finished" completion marker
    master_port: int = 29500
    profile: bool = False

    # ------------------------------------------------------------------ helpers
    def replace(self, **kw) -> "DBSConfig":
        return dataclasses.replace(self, **kw)

    @property
    def is_lm(self) -> bool:
        return self.model == "transformer"

    @property
    def num_classes(self) -> int:
        return 100 if self.dataset == "cifar100" else 10

    def device_for_rank(self, rank: int) -> str:
        """debug → cpu; ``-gpu N`` → every rank on GPU N; ``-gpu a,b,c`` → rank→GPU map
        (reference ``dbs.py:63-71`` and ``dbs.py:517-520``)."""
        if self.debug:
            return "cpu"
        if isinstance(self.gpu, int):
            return f"cuda:{self.gpu}"
        return f"cuda:{self.gpu[rank % len(self.gpu)]}"

    def experiment_id(self, node: Union[int, str] = 0) -> str:
        """Same artifact stem as the reference (``dbs.py:54-61``): e.g.
        ``transformer-wikitext2-debug1-n2-bs64-lr0.0100-ep3-dbs1-ft0-ftc0.100000-node0-ocp0``;
        prefixed ``puredbs=`` under ``-de``."""
        stem = "%s-%s-debug%d-n%d-bs%d-lr%.4f-ep%d-dbs%d-ft%d-ftc%f-node%s-ocp%d" % (
            self.model, self.dataset, int(self.debug), self.world_size, self.batch_size,
            self.learning_rate, self.epoch_size, int(self.dynamic_batch_size),
            int(self.fault_tolerance), self.fault_tolerance_chance, str(node),
            int(self.one_cycle_policy))
        if self.disable_enhancements:
            stem = "puredbs=" + stem
        return stem

    def resolved_dtype(self, device: str) -> str:
        if self.dtype != "auto":
            return self.dtype
        return "bf16" if device.startswith("cuda") else "fp32"

    def resolved_comm(self, device: str) -> str:
        if self.comm != "auto":
            return self.comm
        return "symm" if device.startswith("cuda") else "gloo"

    def resolved_dbs_model(self) -> str:
        if self.dbs_model != "auto":
            return self.dbs_model
        return "proportional" if self.debug else "affine"

    def resolved_clip(self) -> float:
        if self.clip_grad_norm >= 0:
            return self.clip_grad_norm
        return 0.25 if self.is_lm else 0.0   # reference dbs.py:274 clips only the LM
