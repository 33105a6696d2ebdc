#!/usr/bin/env python
"""Evidence from a torch.profiler / Kineto Chrome trace (``dbs.py --profile true`` writes ``<log_dir>/<id>.trace.json``):

  * which kernels ran, per CUDA stream, with their total device time;
  * any ``nccl*`` kernel in the captured steps (the product's gradient path must show none);
  * how much of the fused ``weighted_allreduce_kernel`` device time ran CONCURRENTLY with compute kernels on other
    streams (the bucket collectives are fired from autograd hooks on a communication stream, overlap = backward compute
    still running while a bucket is being reduced).

    python tools/trace_overlap.py logs/<id>.trace.json [--top 15]
"""
import argparse
import collections
import json


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("trace")
    ap.add_argument("--top", type=int, default=15)
    a = ap.parse_args()
    with open(a.trace) as f:
        ev = json.load(f)
    ev = ev["traceEvents"] if isinstance(ev, dict) else ev
    ker = [e for e in ev if e.get("cat") in ("kernel", "Kernel") and "dur" in e]
    if not ker:
        print("no kernel events in trace")
        return
    by_stream = collections.defaultdict(list)
    for e in ker:
        by_stream[e.get("args", {}).get("stream", e.get("tid"))].append(e)
    t0 = min(e["ts"] for e in ker)
    t1 = max(e["ts"] + e["dur"] for e in ker)
    print(f"{len(ker)} kernel events on {len(by_stream)} streams over {1e-3 * (t1 - t0):.2f} ms")
    for st, es in sorted(by_stream.items(), key=lambda kv: -sum(e['dur'] for e in kv[1])):
        tot = sum(e["dur"] for e in es)
        names = collections.Counter()
        for e in es:
            names[e["name"].split("<")[0].split("(")[0][-60:]] += e["dur"]
        top = ", ".join(f"{n} {1e-3 * d:.2f}ms" for n, d in names.most_common(3))
        print(f"  stream {st}: {len(es)} kernels, {1e-3 * tot:.2f} ms busy  [{top}]")
    nccl = [e for e in ker if "nccl" in e["name"].lower()]
    print(f"nccl kernels in trace: {len(nccl)}" + ("" if not nccl else "  e.g. " + nccl[0]["name"][:80]))
    comm = [e for e in ker if "weighted_allreduce" in e["name"]]
    if not comm:
        print("no weighted_allreduce_kernel in trace (single rank?)")
        return
    comm_streams = {e.get("args", {}).get("stream", e.get("tid")) for e in comm}
    other = sorted(((e["ts"], e["ts"] + e["dur"]) for e in ker
                    if e.get("args", {}).get("stream", e.get("tid")) not in comm_streams))
    # merge compute intervals
    merged = []
    for s, t in other:
        if merged and s <= merged[-1][1]:
            merged[-1][1] = max(merged[-1][1], t)
        else:
            merged.append([s, t])
    tot = ov = 0.0
    for e in comm:
        s, t = e["ts"], e["ts"] + e["dur"]
        tot += t - s
        for ms, mt in merged:
            if mt <= s:
                continue
            if ms >= t:
                break
            ov += min(t, mt) - max(s, ms)
    names = collections.Counter(e["name"].split("(")[0][-70:] for e in comm)
    print(f"weighted_allreduce_kernel: {len(comm)} launches on stream(s) {sorted(comm_streams)}, {1e-3 * tot:.3f} ms device time, "
          f"{100.0 * ov / max(tot, 1e-9):.1f}% of it concurrent with compute kernels on other streams")
    for n, c in names.most_common(4):
        print(f"    {c:4d} x {n}")


if __name__ == "__main__":
    main()
