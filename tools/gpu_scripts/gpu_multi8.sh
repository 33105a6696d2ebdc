#!/bin/bash
# 8-GPU box: collectives test, allreduce sweep, headline bench both arms.
mkdir -p gpurun_out
N=${1:-8}
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
echo "== multi-gpu tests"; timeout 900 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -30
echo "== allreduce sweep N=$N"; NCCL_DEBUG=WARN timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29901 tools/allreduce_sweep.py 2> gpurun_out/sweep.err | tail -14; tail -3 gpurun_out/sweep.err
echo "== bench ours N=$N"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29902 bench.py --gpus $N --steps 20 --warmup 5 2> gpurun_out/b$N.err | tee gpurun_out/bench_ours_n$N.json; tail -4 gpurun_out/b$N.err
echo "== bench ours N=4"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29903 bench.py --gpus 4 --steps 20 --warmup 5 2> gpurun_out/b4.err | tee gpurun_out/bench_ours_n4.json | cut -c1-600
echo "== bench ref N=$N"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29904 bench.py --impl reference --gpus $N --steps 20 --warmup 5 2> gpurun_out/r$N.err | tee gpurun_out/bench_ref_n$N.json; tail -4 gpurun_out/r$N.err
echo "== bench ours N=$N no-dbs"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29905 bench.py --gpus $N --steps 20 --warmup 5 --no-dbs 2> gpurun_out/b${N}nodbs.err | tee gpurun_out/bench_ours_n${N}_nodbs.json | cut -c1-700
