#!/bin/bash
mkdir -p gpurun_out
echo "== gemm tests"; timeout 600 python -m pytest tests/test_gpu_gemm.py -x -q 2>&1 | tail -3
echo "== gemm bench"; timeout 300 python tools/bench_gemm.py 2>&1 | tail -4
echo "== multi-gpu tests"; timeout 900 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -15
for flag in "" "--no-overlap"; do
echo "== bench ours N=2 $flag"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29811 bench.py --gpus 2 --steps 20 --warmup 3 $flag 2> gpurun_out/b2.err | tee gpurun_out/bench_ours_n2$flag.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['config']['local_batches'], d['straggler_wait_ms_per_step'])"; tail -2 gpurun_out/b2.err
done
echo "== bench ours N=1"; timeout 600 python bench.py --steps 10 --warmup 3 2>/dev/null | cut -c1-200
