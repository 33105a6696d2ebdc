#!/bin/bash
mkdir -p gpurun_out
echo "== tests"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for v in 1 0; do
echo "== bench densenet conv3=$v"; DLB_TC_CONV3=$v timeout 600 python bench.py --steps 10 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['gpu_launches']/d['steps'])"
done
echo "== bench densenet b128"; timeout 600 python bench.py --steps 20 --warmup 3 --batch 128 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ms/step', d['ms_per_step'], 'launches/step', d['gpu_launches']/d['steps'])"
echo "== resnet50 B=1024 ours"; timeout 600 python bench.py --model resnet50 --batch 1024 --steps 10 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'])"
