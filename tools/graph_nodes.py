#!/usr/bin/env python
"""Summarise a CUDA-graph dot dump (``DLB_GRAPH_DUMP=<prefix> python bench.py ...`` → ``<prefix>.b<B>.rank<r>.dot``,
written by ``cudaGraphDebugDotPrint`` through ``torch.cuda.CUDAGraph.debug_dump``):

* node count per kind (kernel / memset / memcpy / event record+wait / ...),
* kernel-name histogram (demangled prefix),
* number of root/leaf nodes and the **critical path** (longest dependency chain, in nodes and in kernel nodes) —
  at small per-rank batches every node costs a few microseconds of launch + dependency latency, so the chain length,
  not the byte count, bounds the step (BASELINE.md §3 note ii).

    python tools/graph_nodes.py step.b64.rank0.dot [--top 25]
"""
import argparse
import collections
import re
import subprocess
import sys

NODE = re.compile(r'^\s*"?([\w.]+)"?\s*\[(.*)\]\s*;?\s*$')
EDGE = re.compile(r'^\s*"?([\w.]+)"?\s*->\s*"?([\w.]+)"?')
LABEL = re.compile(r'label\s*=\s*"((?:[^"\\]|\\.)*)"')
KINDS = ("KERNEL", "MEMSET", "MEMCPY", "EVENT_RECORD", "EVENT_WAIT", "EVT_RECORD", "EVT_WAIT", "HOST", "MEM_ALLOC", "MEM_FREE",
         "CHILD", "EMPTY", "CONDITIONAL", "BATCH_MEM_OP", "EXT_SEMAS")


def classify(label: str):
    """-> (kind, name).  The label's lines hold the node id, the node kind or (for kernels) the function name."""
    lines = [l.strip() for l in label.replace("\\l", "\\n").split("\\n") if l.strip()]
    up = label.upper()
    for k in KINDS:
        if k in up and k != "KERNEL":
            return k.replace("EVT_", "EVENT_"), k
    name = next((l for l in lines if not l.isdigit() and not l.upper().startswith(("NODE", "ID"))), lines[0] if lines else "?")
    return "KERNEL", name


def demangle(names):
    mangled = [n for n in names if n.startswith("_Z")]
    out = {}
    if mangled:
        try:
            res = subprocess.run(["c++filt"], input="\n".join(mangled), capture_output=True, text=True, timeout=30).stdout.split("\n")
            out = dict(zip(mangled, res))
        except Exception:
            pass
    return out


def short(name: str) -> str:
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    name = re.sub(r"^void ", "", name)
    return re.split(r"[(<]", name)[0][:70]


def parse(path):
    nodes, edges = {}, []
    for line in open(path, errors="replace"):
        m = EDGE.match(line)
        if m:
            edges.append((m.group(1), m.group(2)))
            continue
        m = NODE.match(line)
        if m and "label" in m.group(2) and not m.group(1) in ("graph", "node", "edge"):
            lab = LABEL.search(m.group(2))
            if lab:
                nodes[m.group(1)] = classify(lab.group(1))
    return nodes, edges


def critical_path(nodes, edges):
    succ, indeg = collections.defaultdict(list), collections.Counter()
    for a, b in edges:
        if a in nodes and b in nodes:
            succ[a].append(b); indeg[b] += 1
    depth = {n: (1, 1 if nodes[n][0] == "KERNEL" else 0) for n in nodes}
    order = collections.deque(n for n in nodes if indeg[n] == 0)
    roots = len(order)
    seen = 0
    while order:
        a = order.popleft(); seen += 1
        for b in succ[a]:
            cand = (depth[a][0] + 1, depth[a][1] + (1 if nodes[b][0] == "KERNEL" else 0))
            if cand > depth[b]:
                depth[b] = cand
            indeg[b] -= 1
            if indeg[b] == 0:
                order.append(b)
    leaves = sum(1 for n in nodes if not succ[n])
    best = max(depth.values()) if depth else (0, 0)
    return roots, leaves, best, seen == len(nodes)


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("dot")
    ap.add_argument("--top", type=int, default=25)
    a = ap.parse_args(argv)
    nodes, edges = parse(a.dot)
    kinds = collections.Counter(k for k, _ in nodes.values())
    print(f"{len(nodes)} nodes, {len(edges)} edges")
    for k, c in kinds.most_common():
        print(f"  {c:6d}  {k}")
    roots, leaves, (path_nodes, path_kernels), acyclic = critical_path(nodes, edges)
    print(f"roots {roots}, leaves {leaves}, critical path {path_nodes} nodes ({path_kernels} kernels)" + ("" if acyclic else "  [cycle?]"))
    names = [n for k, n in nodes.values() if k == "KERNEL"]
    dm = demangle(set(names))
    hist = collections.Counter(short(dm.get(n, n)) for n in names)
    print(f"kernel nodes by name (top {a.top}):")
    for n, c in hist.most_common(a.top):
        print(f"  {c:6d}  {n}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
