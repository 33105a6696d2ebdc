#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel family."""
import collections
import csv
import re
import sys


def key(n):
    n = n.replace('void ', '').replace('(anonymous namespace)::', '')
    n = re.sub(r'\(.*$', '', n)
    out, d = '', 0
    for ch in n:
        if ch == '<':
            d += 1
        if d == 0:
            out += ch
        if ch == '>':
            d -= 1
    return out[:90]


def main(path, top=30):
    lines = [l for l in open(path) if not l.startswith('==')]
    rows = list(csv.DictReader(lines))
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in rows:
        v = float(row['Metric Value'].replace(',', ''))
        v *= {'ns': 1, 'us': 1e3, 'ms': 1e6}.get(row['Metric Unit'], 1)
        k = key(row['Kernel Name'])
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"{len(rows)} launches, total {tot / 1e6:.2f} ms (serialised, cold cache: compare shares)")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{v[1] / 1e6:9.2f} ms {100 * v[1] / tot:5.1f}% n={v[0]:5d} avg {v[1] / v[0] / 1e3:8.1f}us  {k}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
