#!/bin/bash
echo "== tests"; timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -k "group_norm or dense_block or densenet" 2>&1 | tail -3
for b in 64 128 512; do
echo "== batch $b NEW"; timeout 600 python bench.py --steps 20 --warmup 3 --batch $b 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'])"
echo "== batch $b PREV"; DLB_NATIVE_LIB=$PWD/baseline/ab/libdlb_prev.so timeout 600 python bench.py --steps 20 --warmup 3 --batch $b 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'])"
done
