#!/bin/bash
mkdir -p gpurun_out
for flag in "" "--no-dbs"; do
echo "== densenet ours N=4 burn $flag"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29811 bench.py --gpus 4 --steps 20 --warmup 5 $flag 2>/dev/null | tee gpurun_out/final_bench_ours_n4$flag.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['config']['local_batches'], d['straggler_wait_ms_per_step'])"
done
echo "== ref N=4"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29812 bench.py --impl reference --gpus 4 --steps 20 --warmup 5 2>/dev/null | tee gpurun_out/final_bench_ref_n4.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'])"
