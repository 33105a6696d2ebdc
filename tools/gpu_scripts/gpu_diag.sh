#!/bin/bash
for b in 64 128 256 512; do
echo "== batch $b"; timeout 600 python bench.py --steps 20 --warmup 3 --batch $b 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'host issue', d['host_issue_ms_per_step'], 'launches/step', d['gpu_launches']/d['steps'])"
done
