#!/bin/bash
# Round 2, 4-GPU call (charged 4x: keep it short): default bench at N=4 (new default balancer), allreduce sweep at 4 GPUs,
# collective tests at 4 GPUs.   gpurun --gpus 4 --timeout 420 -- bash tools/gpu_scripts/r2_n4.sh
N=4
mkdir -p gpurun_out; O=gpurun_out/r2_n4; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== bench ours (default: tf32 headline, bf16 alt, --dbs_model auto, 3 rounds)"
timeout 300 $TR --master-port 29803 bench.py --gpus $N --steps 20 --warmup 5 2> $O/ours.err | tee $O/ours.json | cut -c1-600
echo "== allreduce sweep (1 KB - 256 MiB)"
SWEEP_MAX_BYTES=$((1<<28)) timeout 200 $TR --master-port 29801 tools/allreduce_sweep.py > $O/sweep.jsonl 2> $O/sweep.err; tail -3 $O/sweep.jsonl | cut -c1-420
cp gpurun_out/allreduce_sweep_n$N.json $O/ 2>/dev/null
echo "== collective tests ($N GPUs)"; timeout 200 python -m pytest tests/test_gpu_multi.py -q -rA 2>&1 | tail -6 | tee $O/pytest_multi.txt
echo "== bench ours, reference's rule (proportional) for comparison"
timeout 200 $TR --master-port 29805 bench.py --gpus $N --steps 20 --warmup 5 --dbs-model proportional --alt-dtype "" 2> $O/ours_prop.err | tee $O/ours_prop.json | cut -c1-400
tail -2 $O/*.err | tail -12
