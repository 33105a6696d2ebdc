#!/bin/bash
mkdir -p gpurun_out
for flag in "" "--no-dbs"; do
echo "== densenet ours N=2 burn $flag"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29811 bench.py --gpus 2 --steps 20 --warmup 5 $flag 2>/dev/null | tee gpurun_out/final_bench_ours_n2$flag.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['config']['local_batches'], d['straggler_wait_ms_per_step'])"
done
