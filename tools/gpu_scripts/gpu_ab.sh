#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do
echo "== bench PREV lib"; DLB_NATIVE_LIB=$PWD/baseline/ab/libdlb_prev.so timeout 600 python bench.py --steps 10 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'])"
echo "== bench NEW lib"; timeout 600 python bench.py --steps 10 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'])"
done
