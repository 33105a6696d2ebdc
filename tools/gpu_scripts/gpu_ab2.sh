#!/bin/bash
mkdir -p gpurun_out
echo "== tests"; timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -k "dense_block or densenet or whole_step" 2>&1 | tail -4
for i in 1 2; do
echo "== bench side stream OFF"; DLB_SIDE_STREAM=0 timeout 600 python bench.py --steps 10 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'])"
echo "== bench side stream ON"; DLB_SIDE_STREAM=1 timeout 600 python bench.py --steps 10 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'])"
done
