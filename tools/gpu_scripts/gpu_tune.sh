#!/bin/bash
for kb in 24 96 192 384; do for b in 128 512; do
echo "== min_kb $kb batch $b"; DLB_GN_MIN_KB=$kb timeout 600 python bench.py --steps 20 --warmup 3 --batch $b 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'])"
done; done
