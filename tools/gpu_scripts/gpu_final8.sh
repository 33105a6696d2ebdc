#!/bin/bash
mkdir -p gpurun_out
N=8
echo "== bench ours N=8"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29902 bench.py --gpus $N --steps 20 --warmup 5 2> gpurun_out/g8.err | tee gpurun_out/final_bench_ours_n8.json; grep -v -i warning gpurun_out/g8.err | tail -3
echo "== bench ref N=8"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29904 bench.py --impl reference --gpus $N --steps 20 --warmup 5 2> gpurun_out/g8r.err | tee gpurun_out/final_bench_ref_n8.json | cut -c1-300
echo "== resnet50 B=1024 ours N=8"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29906 bench.py --gpus $N --model resnet50 --batch 1024 --steps 20 --warmup 5 2> gpurun_out/g8r50.err | tee gpurun_out/final_bench_ours_resnet50_n8.json | cut -c1-700
echo "== transformer B=512 ours N=8"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29907 bench.py --gpus $N --model transformer --batch 512 --steps 20 --warmup 5 2> gpurun_out/g8lm.err | tee gpurun_out/final_bench_ours_transformer_n8.json | cut -c1-700; grep -v -i warning gpurun_out/g8lm.err | tail -3
echo "== sweep (large sizes)"; SWEEP_MAX_BYTES=268435456 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29901 tools/allreduce_sweep.py 2>/dev/null | tail -5 | cut -c1-420
