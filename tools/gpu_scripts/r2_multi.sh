#!/bin/bash
# Multi-GPU evidence, N = $1 GPUs of one box:   gpurun --gpus N --timeout 1500 -- bash tools/gpu_scripts/r2_multi.sh N [quick]
#   collective numerics at N GPUs, allreduce bus-bandwidth sweep 1 KB-1 GB vs scale+NCCL, both bench arms, DBS models A/B,
#   a profiler trace of graph-replayed steps (stream overlap, absence of nccl kernels), compute-sanitizer on the collectives.
N=${1:-2}; QUICK=${2:-}
mkdir -p gpurun_out; O=gpurun_out/r2_multi_n$N; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi topo -m > $O/topo.txt 2>&1
echo "== collective tests ($N GPUs)"; timeout 600 python -m pytest tests/test_gpu_multi.py -q -rA 2>&1 | tail -15 | tee $O/pytest_multi.txt
echo "== allreduce sweep";  SWEEP_MAX_BYTES=$((1<<30)) timeout 600 $TR --master-port 29801 tools/allreduce_sweep.py > $O/sweep.jsonl 2> $O/sweep.err; tail -4 $O/sweep.jsonl | cut -c1-400
cp gpurun_out/allreduce_sweep_n$N.json $O/ 2>/dev/null
echo "== bench reference";  timeout 900 $TR --master-port 29802 bench.py --impl reference --gpus $N --steps 20 --warmup 5 2> $O/ref.err | tee $O/ref.json | cut -c1-400
echo "== bench ours (default)"; timeout 600 $TR --master-port 29803 bench.py --gpus $N --steps 20 --warmup 5 2> $O/ours.err | tee $O/ours.json | cut -c1-500
if [ -z "$QUICK" ]; then
echo "== bench ours --dbs-model affine (3 rounds)"; timeout 600 $TR --master-port 29804 bench.py --gpus $N --steps 20 --warmup 5 --dbs-model affine --dbs-rounds 3 --alt-dtype "" 2> $O/ours_affine.err | tee $O/ours_affine.json | cut -c1-500
echo "== bench ours --dbs-model proportional (3 rounds)"; timeout 600 $TR --master-port 29805 bench.py --gpus $N --steps 20 --warmup 5 --dbs-model proportional --dbs-rounds 3 --alt-dtype "" 2> $O/ours_prop3.err | tee $O/ours_prop3.json | cut -c1-500
echo "== bench ours --no-dbs"; timeout 600 $TR --master-port 29806 bench.py --gpus $N --steps 20 --warmup 5 --no-dbs --alt-dtype "" 2> $O/ours_nodbs.err | tee $O/ours_nodbs.json | cut -c1-500
fi
if [ "$N" = "8" ]; then
  echo "== config #3: ResNet-50, B=1024, uniform ranks (DBS must stay at the equal split)"
  timeout 900 $TR --master-port 29812 bench.py --impl reference --gpus $N --steps 20 --warmup 5 --model resnet50 --batch 1024 --throttle-ms 0 2> $O/r50_ref.err | tee $O/r50_ref.json | cut -c1-300
  timeout 600 $TR --master-port 29813 bench.py --gpus $N --steps 20 --warmup 5 --model resnet50 --batch 1024 --throttle-ms 0 2> $O/r50_ours.err | tee $O/r50_ours.json | cut -c1-400
  echo "== config #4: Transformer LM (wikitext-2 shape), DBS, bf16"
  timeout 900 $TR --master-port 29814 bench.py --impl reference --gpus $N --steps 20 --warmup 5 --model transformer --batch 512 2> $O/lm_ref.err | tee $O/lm_ref.json | cut -c1-300
  timeout 600 $TR --master-port 29815 bench.py --gpus $N --steps 20 --warmup 5 --model transformer --batch 512 --dtype bf16 2> $O/lm_ours.err | tee $O/lm_ours.json | cut -c1-400
fi
echo "== profiler trace of graph-replayed steps (rank 0)"
DLB_PROFILE_GRAPHS=1 timeout 600 $TR --master-port 29808 dbs.py -d false -ws $N -b 512 -m densenet -ds cifar10 -e 1 --synthetic true \
    --train_samples 10240 --test_samples 256 --validate false --profile true --throttle_rank $((N-1)) --throttle_ms 3 --throttle_mode burn \
    --log_dir $O/logs --stats_dir $O/statis --force true > $O/profile.out 2> $O/profile.err
T=$(ls $O/logs/*node0*.trace.json 2>/dev/null | head -1)
[ -n "$T" ] && python tools/trace_overlap.py $T | tee $O/trace_overlap.txt && rm -f $O/logs/*.trace.json
if [ "$N" = "2" ]; then
  echo "== README headline mode: 4 workers on 2 GPUs (-gpu 0,0,0,1), several ranks share GPU 0 (reference README.md:23-29)"
  (cd $O && timeout 900 python ../../dbs.py -d false -ws 4 -b 512 -m densenet -ds cifar10 -gpu 0,0,0,1 -e 3 --synthetic true \
      --train_samples 10240 --test_samples 256 --validate false --master_port 29811 --force true > oversub.out 2> oversub.err; \
   python - <<'PY'
import glob, numpy as np
for f in glob.glob("statis/*.npy"):
    d = np.load(f, allow_pickle=True).item()
    print("oversubscription -gpu 0,0,0,1: local batches per epoch", d["local_batches"], "samples/s", [round(x) for x in d["samples_per_sec"]])
PY
  ) 2>&1 | tail -3 | tee $O/oversub.txt
  echo "== compute-sanitizer on the collectives"
  bash tools/gpu_scripts/sanitize.sh comm-only 2>&1 | tail -6 | tee $O/sanitize.txt
  cp gpurun_out/sanitize_comm_* $O/ 2>/dev/null
fi
tail -2 $O/*.err | tail -40
