#!/bin/bash
echo "== tests (PDL on)"; DLB_PDL=1 timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_gemm.py -x -q 2>&1 | tail -3
for b in 128 512; do for v in 0 1; do
echo "== batch $b PDL=$v"; DLB_PDL=$v timeout 600 python bench.py --steps 20 --warmup 3 --batch $b 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'])"
done; done
