"""Training engine: the DBS control loop.

Per epoch (capabilities of reference ``dbs.py:313-446``; call stack in SURVEY §3.2-3.4):

    LR policy → reallocator.step() (throughput-proportional re-split of the global batch)
    → re-partition data so every rank runs the SAME number of steps with ITS local batch
    → train (fwd, bwd, straggler injection, weighted gradient allreduce, SGD)
    → validate → exchange per-rank pure-compute time → feed the reallocator → record.

What is different by design:
  * one flat parameter/gradient store and a fused  pack(+w_r,+clip) → allreduce → SGD  pipeline
    (``parallel/buckets.py``) instead of one scale + one blocking allreduce per parameter tensor;
  * on CUDA the whole step (augment → forward → backward → pack → collective → SGD) is captured in a
    CUDA graph per local batch size and replayed; the graph is re-captured when DBS changes the size;
  * time accounting is device-side (CUDA events + in-kernel barrier-wait counter), never
    ``time.time()`` around asynchronous launches;
  * datasets / corpus are built once, only indices are re-sliced per epoch (D11); the loss is
    accumulated on the device and read back every ``log_every`` steps instead of two ``.item()``
    syncs per step (K22).
"""
from __future__ import annotations

import math
import os
import time
from typing import Dict, Tuple

import numpy as np
import torch
import torch.distributed as dist
import torch.nn.functional as F

from .. import ops
from ..balance import AffineReallocator, Reallocator, TimeTracker
from ..config import DBSConfig
from ..data import (BatchStager, DataPartitioner, batchify, load_corpus, load_image_dataset, split_token_stream)
from ..fault import StragglerInjector
from ..models import build_model
from ..parallel import FlatState, make_comm
from ..utils import StatsRecorder, Tracer, load_checkpoint, save_checkpoint
from .lr_policy import lr_at_epoch

# tf32 = fp32 storage with TF32 tensor-core math in the conv / linear kernels (tcgen05.mma kind::tf32): the precision class of
# the reference's default PyTorch path (fp32 model, dbs.py:363; cuDNN convolutions run TF32 by default).  "fp32" is the same
# storage; on CUDA its convolutions take the same TF32 kernels unless DLB_TF32=0 (then the vendor library decides).
_DT = {"fp32": torch.float32, "bf16": torch.bfloat16, "tf32": torch.float32}


class Trainer:
    def __init__(self, cfg: DBSConfig, rank: int, world: int, device: str, logger, comm=None):
        self.cfg, self.rank, self.world, self.logger = cfg, rank, world, logger
        self.device = torch.device(device)
        self.cuda = self.device.type == "cuda"
        if self.cuda:
            torch.cuda.set_device(self.device)
        self.dtype = _DT[cfg.resolved_dtype(device)]
        self.dtype_name = cfg.resolved_dtype(device)
        if self.dtype_name == "fp32" and self.cuda and os.environ.get("DLB_TF32", "1") == "1":
            self.dtype_name = "tf32"
        self.comm = comm if comm is not None else make_comm(cfg.resolved_comm(device), self.device, algo=cfg.allreduce_algo,
                                                            timeout_s=cfg.comm_timeout_s)
        self.is_lm = cfg.is_lm
        self.log_every = 10
        self.max_cached_graphs = 4
        self._host_sleep_s = 0.0
        self._graphs: Dict[int, "GraphedStep"] = {}
        self._eager_steps_at: Dict[int, int] = {}
        # --profile: NVTX ranges + a torch.profiler capture of a few steps + host phase table (utils/tracing.py)
        self.tracer = Tracer(cfg.profile, os.path.join(cfg.log_dir, cfg.experiment_id(rank)) if cfg.profile else None,
                             self.cuda, logger=logger)
        self._build()

    # ------------------------------------------------------------------------------------------------
    def _build(self) -> None:
        cfg = self.cfg
        # ---- data (built once) ----
        if self.is_lm:
            sizes = None
            if cfg.train_samples:
                sizes = {"train": cfg.train_samples, "valid": max(1000, cfg.train_samples // 10),
                         "test": cfg.test_samples or max(1000, cfg.train_samples // 10)}
            self.corpus = load_corpus(cfg.corpus_root, cfg.synthetic, sizes, cfg.seed)
            self.ntokens = self.corpus.ntokens
            self.train_set = self.test_set = None
        else:
            self.train_set = load_image_dataset(cfg.dataset, True, cfg.data_root, cfg.synthetic, cfg.train_samples,
                                                cfg.seed, pin=self.cuda)
            self.test_set = load_image_dataset(cfg.dataset, False, cfg.data_root, cfg.synthetic, cfg.test_samples,
                                               cfg.seed, pin=self.cuda)
            self.ntokens = 0
        # ---- model: identical init on every rank, then per-rank dropout streams (fixes D20) ----
        torch.manual_seed(cfg.seed)
        self.model = build_model(cfg.model, cfg.num_classes, self.ntokens)
        self.flat = FlatState(self.model, self.device, self.dtype, self.comm, cfg.learning_rate, cfg.momentum,
                              0.0, cfg.bucket_mb, _DT[cfg.wire_dtype], cfg.resolved_clip(), cfg.clip_mode,
                              seed_weighting=self.cuda and not self.is_lm)
        self.flat.sync_initial_params()
        if cfg.overlap_comm:
            self.flat.enable_overlap()
        torch.manual_seed(cfg.seed + 1 + self.rank)
        if self.cuda:
            torch.cuda.manual_seed(cfg.seed + 1 + self.rank)
        # ---- balancer, injector, recorder ----
        realloc_cls = AffineReallocator if cfg.resolved_dbs_model() == "affine" else Reallocator
        self.realloc = realloc_cls(self.world, cfg.batch_size, cfg.dynamic_batch_size, cfg.rounding,
                                   cfg.min_local_batch, cfg.batch_quantum, cfg.time_ema)
        self.injector = StragglerInjector(self.rank, cfg.fault_tolerance, cfg.fault_tolerance_chance,
                                          cfg.throttle_rank, cfg.throttle_ms, cfg.throttle_mode, self.device,
                                          logger=self.logger)
        self.recorder = StatsRecorder(cfg, enabled=(self.rank == 0))
        self.tracker = TimeTracker(self.device)
        self.stager = None
        if not self.is_lm:
            self.stager = BatchStager(self.train_set, cfg.batch_size, self.device)
        self.start_epoch = 0
        if cfg.resume and cfg.checkpoint_dir:
            blob = load_checkpoint(cfg, self.flat, self.realloc)
            if blob is not None:
                self.start_epoch = int(blob["epoch"]) + 1
                self._resume_extra = blob.get("extra") or {}
                self.logger.info(f"Rank {self.rank} resumed from epoch {blob['epoch']}")
        self.total_train_time = 0.0
        self.global_step = 0
        self.step_t = torch.zeros(1, dtype=torch.int64, device=self.device)
        extra = getattr(self, "_resume_extra", None)
        if extra:                      # counters that seed the augmentation stream + the stats history of the resumed run
            self.global_step = int(extra.get("global_step", 0))
            self.step_t.fill_(int(extra.get("step_t", 0)))
            self.total_train_time = float(extra.get("total_train_time", 0.0))
            if extra.get("recorder"):
                self.recorder.load_state(extra["recorder"])
        self.ts = torch.zeros(2, dtype=torch.int64, device=self.device)      # [0] step-start stamp, [1] accumulated compute ns
        self._dev_timers = self.cuda and ops._native.available() and hasattr(ops._native.get(), "dlb_stamp")
        self.loss_acc = torch.zeros(1, dtype=torch.float32, device=self.device)

    # ------------------------------------------------------------------------------------------------
    # one optimisation step, eager flavour (CPU, or CUDA before a graph exists for this batch size)
    def _stamp_start(self) -> None:
        if self._dev_timers:
            nat = ops._native
            nat.check(nat.get().dlb_stamp(self.ts.data_ptr(), nat.stream_ptr(self.device)), "stamp")

    def _stamp_compute_end(self) -> None:
        """End of this rank's compute on the main stream (before it joins the communication stream)."""
        if self._dev_timers:
            nat = ops._native
            nat.check(nat.get().dlb_stamp_acc(self.ts.data_ptr(), self.ts.data_ptr() + 8, nat.stream_ptr(self.device)), "stamp_acc")

    def _forward_backward(self, x, y) -> torch.Tensor:
        with self.tracer.range("forward"):
            if self.is_lm:
                loss = self.model.forward_loss(x, y)
            else:
                out = self.model(x)
                loss = ops.cross_entropy(out, y, grad_scale=self.flat.seed_scale())
        with self.tracer.range("backward"):
            loss.backward()
        return loss.detach()

    def _prepare_images(self, imgs_u8: torch.Tensor) -> torch.Tensor:
        ds = self.train_set
        return ops.augment(imgs_u8, ds.mean, ds.std, ds.pad, ds.flip, self.cfg.seed + self.rank, self.global_step,
                           self.dtype, step_tensor=self.step_t if (self.cuda and ops._native.available()) else None)

    def _eager_step(self, xb, yb, steady: bool = True) -> torch.Tensor:
        self.model.train()
        self.tracker.start_compute()
        if steady:
            self._stamp_start()
        with self.tracer.range("augment"):
            x = xb if self.is_lm else self._prepare_images(xb)
        loss = self._forward_backward(x, yb)
        with self.tracer.range("straggler"):
            self.injector.device_delay()               # between backward and allreduce (reference dbs.py:236)
        if steady:
            self._stamp_compute_end()
        self.tracker.stop_compute(steady)
        # the host sleep sits OUTSIDE the device-stamped region (still between backward and the allreduce) and is booked
        # explicitly: inside it, a host-bound eager step would count the sleep once in the stamp delta and once here
        slept = self.injector.host_delay()
        if slept and (steady or not self.cuda):
            self.tracker.add_compute(slept)
            if self.cuda:
                self._host_sleep_s += slept
        with self.tracer.range("reduce_and_step"):
            waited = self.flat.reduce_and_step(self.rank)
            self.flat.zero_grad()
        self.tracker.add_sync(waited)
        self.loss_acc += loss.float()
        if self.cuda:
            self.step_t += 1
        return loss

    def train_step(self, xb, yb) -> None:
        """Dispatch: CUDA-graph replay when enabled and warmed up for this batch size, else eager."""
        b = int(yb.shape[0]) if not self.is_lm else int(xb.shape[1])
        # gloo collectives (several ranks sharing a GPU, reference `-gpu 0,0,0,1`) cannot be stream-captured
        # profiling runs stay eager: a replayed graph is one opaque launch, the NVTX ranges would be empty
        # (DLB_PROFILE_GRAPHS=1 keeps the graph under --profile: CUPTI still records the replayed kernels with their streams,
        # which is what tools/trace_overlap.py needs to show the bucket collectives overlapping the backward pass)
        use_graph = (self.cuda and self.cfg.cuda_graphs and ops._native.available() and self.comm.name != "gloo"
                     and not (self.cfg.profile and os.environ.get("DLB_PROFILE_GRAPHS", "0") != "1"))
        if use_graph:
            g = self._graphs.get(b)
            if g is None:
                n = self._eager_steps_at.get(b, 0)
                if n < 3:                                   # warm-up steps at a new size run eagerly
                    self._eager_steps_at[b] = n + 1
                    self._eager_step(xb, yb, steady=False)
                    self.comm.device_wait_seconds(reset=True)      # waits of warm-up steps are not part of the signal either
                    self.global_step += 1
                    return
                from .graph_step import GraphedStep
                while len(self._graphs) >= self.max_cached_graphs:      # DBS keeps moving the local batch: bound the
                    old = next(iter(self._graphs))                      # number of captured graphs (each owns a memory pool)
                    del self._graphs[old]
                    self._eager_steps_at.pop(old, None)
                g = GraphedStep(self, xb, yb)
                self._graphs[b] = g
            slept = self.injector.host_delay()
            if slept:
                self.tracker.add_compute(slept)
                self._host_sleep_s += slept
            self.tracker.start_compute()
            g.replay(xb, yb)
            self.tracker.stop_compute()
        else:
            self._eager_step(xb, yb)
        self.global_step += 1
        self.tracer.step()

    # ------------------------------------------------------------------------------------------------
    def _train_epoch_vision(self, epoch: int, shard, steps: int) -> Tuple[float, float, float]:
        cfg = self.cfg
        g = torch.Generator().manual_seed(cfg.seed * 7919 + epoch * 131 + self.rank)
        order = torch.randperm(len(shard), generator=g).numpy()          # DataLoader(shuffle=True)
        self.comm.barrier()
        self.reset_timers()
        self.loss_acc.zero_()
        running_mark = 0.0
        for step in range(steps):
            with self.tracer.range("stage_h2d"):
                idx = shard.batch_indices(step, order)
                xb, yb = self.stager.stage(idx)
            self.train_step(xb, yb)
            self.stager.release()
            if step % self.log_every == 0 and step > 0:
                acc = float(self.loss_acc.item())
                self.logger.info(f"Rank {self.rank}, epoch {epoch}: {step}, train_loss {(acc - running_mark) / self.log_every}")
                running_mark = acc
        return self._finish_epoch(epoch, steps)

    def _train_epoch_lm(self, epoch: int, tokens: torch.Tensor, local_batch: int, steps: int):
        cfg = self.cfg
        data = batchify(tokens, local_batch)                     # [rows, b]
        if self.cuda:
            data = data.pin_memory()
        self.comm.barrier()
        self.reset_timers()
        self.loss_acc.zero_()
        running_mark = 0.0
        for step in range(steps):
            i = step * cfg.bptt
            src = data[i:i + cfg.bptt]
            tgt = data[i + 1:i + 1 + cfg.bptt].reshape(-1)
            with self.tracer.range("stage_h2d"):
                if self.cuda:
                    src = src.to(self.device, non_blocking=True)
                    tgt = tgt.to(self.device, non_blocking=True)
            self.train_step(src, tgt)
            if step % self.log_every == 0 and step > 0:
                acc = float(self.loss_acc.item())
                self.logger.info(f"Rank {self.rank}, epoch {epoch}: {step}, train_loss {(acc - running_mark) / self.log_every}")
                running_mark = acc
        return self._finish_epoch(epoch, steps)

    def _finish_epoch(self, epoch: int, steps: int):
        compute_s, sync_s, wall_s = self.epoch_times()
        loss = float(self.loss_acc.item()) / max(1, steps)
        self.comm.check_errors()
        self.logger.info(f"Rank {self.rank}, epoch {epoch}, train_time {wall_s}, train_loss {loss}")
        return compute_s, sync_s, loss, wall_s

    def reset_timers(self) -> None:
        self.tracker.reset()
        self._host_sleep_s = 0.0
        if self.cuda:
            self.ts.zero_()
            self.comm.device_wait_seconds(reset=True)

    def epoch_times(self):
        """(compute_s, sync_s, wall_s) of the steps since the last tracker reset.

        compute = this rank's own work per step (augment -> forward -> backward -> injected straggle), taken from the
        device-side stamps that bracket it on the main stream; sync = the rest of the step (waiting for peers inside the
        collectives + the optimizer).  The reference uses the same split (compute = wall − Σ wait, dbs.py:250) but with
        host clocks around asynchronous launches.  Only steady steps are measured; totals are extrapolated."""
        steady, unsteady = self.tracker.steps, self.tracker.unsteady
        total_s, host_sync_s, wall_s = self.tracker.finish()
        scale = (steady + unsteady) / steady if (steady and unsteady) else 1.0
        if self._dev_timers and (self._graphs or self.cuda):
            if steady == 0 and unsteady:
                # every step of this epoch was an unmeasured warm-up step at a new local batch: fall back to the event-timed
                # eager steps instead of reporting ~0 (which the reallocator would read as an infinitely fast rank)
                return max(1e-6, self.tracker.unsteady_compute_s + self._host_sleep_s), host_sync_s, wall_s
            compute_s = float(self.ts[1].item()) * 1e-9 * scale + self._host_sleep_s
            sync_s = max(0.0, total_s - compute_s) + host_sync_s
            if not self._graphs:
                # eager steps time only the compute region with events; the collective's barrier wait comes from the
                # in-kernel counter instead
                sync_s += self.comm.device_wait_seconds() * scale
            return max(1e-6, compute_s), sync_s, wall_s
        return total_s, host_sync_s, wall_s

    # ------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def validate(self, epoch: int) -> Tuple[float, float]:
        """Sharded evaluation (every rank scores 1/world of the test set; the reference scores the whole
        set on every rank and double-normalises the loss).  Returns (mean loss, accuracy %
        for vision | perplexity for the LM)."""
        self.model.eval()
        stats = torch.zeros(3, dtype=torch.float64, device=self.device)      # loss_sum, correct, count
        if self.is_lm:
            data = batchify(self.corpus.test, 10)                             # eval batch 10 (dataloader.py:109)
            rows = list(range(0, data.size(0) - 1, self.cfg.bptt))
            for k, i in enumerate(rows):
                if k % self.world != self.rank:
                    continue
                seq = min(self.cfg.bptt, data.size(0) - 1 - i)
                src = data[i:i + seq].to(self.device)
                tgt = data[i + 1:i + 1 + seq].reshape(-1).to(self.device)
                logp = self.model(src).reshape(-1, self.ntokens)
                stats[0] += F.nll_loss(logp.float(), tgt, reduction="sum").double()
                stats[2] += tgt.numel()
        else:
            ds = self.test_set
            n = len(ds)
            bs = 500
            for k, s in enumerate(range(0, n, bs)):
                if k % self.world != self.rank:
                    continue
                imgs = ds.images[s:s + bs].to(self.device, non_blocking=True)
                labels = ds.labels[s:s + bs].to(self.device, non_blocking=True)
                x = ops.augment(imgs, ds.mean, ds.std, 0, False, 0, 0, self.dtype)
                out = self.model(x).float()
                stats[0] += F.cross_entropy(out, labels, reduction="sum").double()
                stats[1] += (out.argmax(1) == labels).sum().double()
                stats[2] += labels.numel()
        if self.world > 1 and dist.is_initialized():
            dist.all_reduce(stats)
        cnt = max(1.0, float(stats[2].item()))
        val_loss = float(stats[0].item()) / cnt
        metric = math.exp(min(val_loss, 50.0)) if self.is_lm else 100.0 * float(stats[1].item()) / cnt
        name = "perplexity" if self.is_lm else "accuracy"
        self.logger.info(f"Rank {self.rank}, epoch {epoch}, val_loss {val_loss}, {name} {metric}")
        self.model.train()
        return val_loss, metric

    # ------------------------------------------------------------------------------------------------
    def run(self) -> StatsRecorder:
        cfg = self.cfg
        self.logger.info(f"Initiating Rank {self.rank}, World Size {self.world}")
        self.logger.info(f"Rank {self.rank} start training")
        for epoch in range(self.start_epoch, cfg.epoch_size):
            lr = lr_at_epoch(cfg.learning_rate, epoch, cfg.epoch_size, cfg.lr_policy, cfg.one_cycle_policy,
                             cfg.disable_enhancements)
            self.flat.set_lr(lr)
            with self.tracer.range("rebalance"):
                fractions, local_batches = self.realloc.step()
            if cfg.dynamic_batch_size:
                self.logger.info(f"Rank {self.rank}, adjusted partition size to {fractions}")
            self.flat.set_weights(self.realloc.weights(uniform=cfg.disable_enhancements))
            # ---- epoch = one segment (reference: re-split once per epoch, dbs.py:388-391) or, with --rebalance_every N,
            # segments of N steps with a time exchange + re-split between them (indices only are re-partitioned) ----------
            B = cfg.batch_size
            if self.is_lm:
                total_steps = (self.corpus.train.numel() // B - 1) // cfg.bptt
            else:
                total_steps = len(self.train_set) // B
            if cfg.max_steps_per_epoch:
                total_steps = min(total_steps, cfg.max_steps_per_epoch)
            seg_cap = cfg.rebalance_every if (cfg.rebalance_every > 0 and cfg.dynamic_batch_size) else 0
            done = used = 0
            agg = {"compute": 0.0, "sync": 0.0, "loss": 0.0, "wall": 0.0}
            while done < total_steps:
                b = int(local_batches[self.rank])
                seg = min(seg_cap, total_steps - done) if seg_cap else total_steps
                if self.is_lm:
                    if seg_cap:
                        rows = seg * cfg.bptt + 1
                        start = used + rows * int(sum(int(x) for x in local_batches[:self.rank]))
                        tokens = self.corpus.train[start:start + rows * b]
                        consumed = seg * cfg.bptt * B
                    else:
                        pieces = split_token_stream(self.corpus.train.numel(), local_batches)
                        tokens = self.corpus.train[pieces[self.rank]]
                        consumed = 0
                    steps, length = seg, (tokens.numel() // b) * b
                else:
                    part = DataPartitioner(len(self.train_set), local_batches, cfg.seed, True, seg, start=used)
                    shard = part.use(self.rank)
                    steps, length, consumed = part.steps, len(shard), part.steps * B
                self.logger.info(f"Rank {self.rank}, number of batches {steps}, batch size {b}, length {length}")
                if done == 0:
                    self.injector.begin_epoch(epoch, total_steps)
                t0 = time.perf_counter()
                if self.is_lm:
                    compute_s, sync_s, loss, wall_s = self._train_epoch_lm(epoch, tokens, b, steps)
                else:
                    compute_s, sync_s, loss, wall_s = self._train_epoch_vision(epoch, shard, steps)
                self.total_train_time += time.perf_counter() - t0
                agg["compute"] += compute_s; agg["sync"] += sync_s; agg["loss"] += loss * steps; agg["wall"] += wall_s
                done += steps
                used += consumed
                # ---- DBS feedback: exchange pure compute time (reference dbs.py:423-426) ----
                nodes_time = self.comm.gather_times(compute_s)
                if cfg.dynamic_batch_size:
                    self.realloc.observe(nodes_time)
                    self.logger.info(f"Rank {self.rank}, total time {nodes_time}")
                if seg_cap and done < total_steps:
                    fractions, local_batches = self.realloc.step()
                    self.logger.info(f"Rank {self.rank}, step {done}: adjusted partition size to {fractions}")
                    self.flat.set_weights(self.realloc.weights(uniform=cfg.disable_enhancements))
            steps = max(1, done)
            compute_s, sync_s, loss, wall_s = agg["compute"], agg["sync"], agg["loss"] / steps, agg["wall"]
            with self.tracer.range("validate"):
                val_loss, metric = self.validate(epoch) if cfg.validate else (float("nan"), float("nan"))
            self.recorder.append(epoch=epoch, train_loss=loss, train_time=compute_s, sync_time=sync_s, val_loss=val_loss,
                                 accuracy=metric, partition=np.asarray(fractions), node_time=list(nodes_time),
                                 wallclock_time=self.total_train_time, local_batches=[int(x) for x in local_batches],
                                 samples_per_sec=(steps * cfg.batch_size / wall_s) if wall_s > 0 else 0.0,
                                 straggler_wait_ms_per_step=1e3 * sync_s / max(1, steps), steps=steps, lr=lr)
            if cfg.checkpoint_dir and self.rank == 0:
                save_checkpoint(cfg, epoch, self.flat, self.realloc,
                                extra={"global_step": self.global_step, "step_t": int(self.step_t.item()),
                                       "total_train_time": self.total_train_time, "recorder": self.recorder.state()})
        if self.rank == 0:
            self.recorder.save()
        self.tracer.close()
        self.logger.info(f"Rank {self.rank} Terminated")
        self.logger.info(f"Rank {self.rank} Total Time:")
        self.logger.info(self.total_train_time)
        return self.recorder

    def close(self) -> None:
        self.tracer.close()
        self._graphs.clear()
        self.comm.close()
