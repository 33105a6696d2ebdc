#!/bin/bash
# 2-GPU box: full GPU test-suite (incl. fused collectives), then both bench arms at N=1 and N=2.
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
echo "== pytest gpu (incl. multi-GPU)"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25
echo "== bench ours N=1"; timeout 600 python bench.py --steps 10 --warmup 3 2> gpurun_out/b1.err | tee gpurun_out/bench_ours_n1.json | cut -c1-330; tail -3 gpurun_out/b1.err
echo "== bench ours N=2"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29811 bench.py --gpus 2 --steps 10 --warmup 3 2> gpurun_out/b2.err | tee gpurun_out/bench_ours_n2.json; tail -5 gpurun_out/b2.err
echo "== bench ref N=2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29812 bench.py --impl reference --gpus 2 --steps 10 --warmup 3 2> gpurun_out/r2.err | tee gpurun_out/bench_ref_n2.json; tail -5 gpurun_out/r2.err
