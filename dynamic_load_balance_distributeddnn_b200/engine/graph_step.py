"""Whole-step CUDA graph: augment → forward → backward → straggler burner → pack(+w_r,+clip) →
fused weighted allreduce → SGD, captured once per local batch size and replayed.

DenseNet-121 at CIFAR resolution is ~1500 small kernels per step; launch latency, not FLOPs, bounds
the eager reference.  Everything a replay needs to vary lives in device memory — the
learning rate, the DBS weight vector, the augmentation step counter, the burner duration — so the
graph never has to be re-captured except when the rebalancer changes this rank's batch size
(SURVEY §7.4 item 4); graphs are cached per size.
"""
from __future__ import annotations

import os

import torch

from .. import ops


class GraphedStep:
    def __init__(self, trainer, xb: torch.Tensor, yb: torch.Tensor):
        self.t = trainer
        self.x = torch.empty_like(xb)
        self.y = torch.empty_like(yb)
        self.x.copy_(xb)
        self.y.copy_(yb)
        # keep_graph: the cudaGraph_t stays alive after capture so its nodes can be counted (dlb_graph_node_counts) -- the
        # number of dependent nodes, not FLOPs, bounds the step at small per-rank batches
        try:
            self.graph = torch.cuda.CUDAGraph(keep_graph=True)
            self._keep = True
        except TypeError:
            self.graph = torch.cuda.CUDAGraph()
            self._keep = False
        self.node_counts = None
        # DLB_GRAPH_DUMP=<prefix>: write the captured graph as <prefix>.b<batch>.rank<r>.dot (cudaGraphDebugDotPrint);
        # `python tools/graph_nodes.py <file>` reports node counts per kind / kernel and the critical-path length --
        # the quantity that bounds the step at small per-rank batches
        dump = os.environ.get("DLB_GRAPH_DUMP")
        if dump:
            self.graph.enable_debug_mode()
        t = trainer
        t.model.train()
        t.flat.zero_grad()
        torch.cuda.synchronize(t.device)
        launches0 = ops._native.launch_count()
        with torch.cuda.graph(self.graph):
            t._stamp_start()
            x = self.x if t.is_lm else t._prepare_images(self.x)
            loss = t._forward_backward(x, self.y)
            t.injector.device_delay()
            t._stamp_compute_end()
            t.flat.reduce_and_step(t.rank)
            t.flat.zero_grad()
            t.loss_acc += loss.float()
            t.step_t += 1
        self.native_launches = ops._native.launch_count() - launches0
        if self._keep:
            try:
                import ctypes
                lib = ops._native.get()
                if lib is not None and hasattr(lib, "dlb_graph_node_counts"):
                    lib.dlb_graph_node_counts.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
                    lib.dlb_graph_node_counts.restype = ctypes.c_int
                    buf = (ctypes.c_longlong * 9)()
                    if lib.dlb_graph_node_counts(ctypes.c_void_p(int(self.graph.raw_cuda_graph())), buf) == 0:
                        keys = ("total", "kernel", "memcpy", "memset", "host", "event_record", "event_wait", "other", "edges")
                        self.node_counts = dict(zip(keys, [int(v) for v in buf]))
                        t.graph_nodes = self.node_counts
            except Exception:          # noqa: BLE001 - introspection only
                self.node_counts = None
            self.graph.instantiate()
        if dump:
            b = int(yb.shape[0]) if not t.is_lm else int(xb.shape[1])
            self.graph.debug_dump(f"{dump}.b{b}.rank{t.rank}.dot")
        t.flat._keep = None

    def replay(self, xb: torch.Tensor, yb: torch.Tensor) -> None:
        self.x.copy_(xb, non_blocking=True)
        self.y.copy_(yb, non_blocking=True)
        self.graph.replay()
        lib = ops._native.get()
        if lib is not None:
            lib.dlb_launch_count_add(self.native_launches)
