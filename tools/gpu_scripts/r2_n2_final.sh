#!/bin/bash
# Round 2, last 2-GPU call (charged 2x: short): default bench at N=2 with the final balancer (pooled-slope affine, typical-rank
# grouping, 3 rounds x 10 steps).
N=2
mkdir -p gpurun_out; O=gpurun_out/r2_n2_final; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 240 $TR --master-port 29803 bench.py --gpus $N --steps 20 --warmup 5 --alt-dtype "" 2> $O/ours.err | tee $O/ours.json | cut -c1-300
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_n2_final/ours.json").read().strip().splitlines()[-1])
a = d.get("alt") or {}
print("N=2", d["value"], d["ms_per_step"], d["dtype"], "| alt", a.get("value"), a.get("ms_per_step"), "| split", d["detail"]["local_batches"],
      d["detail"]["dbs_model"], "wait", d.get("straggler_wait_ms_per_step"))
PY
grep "dbs round" $O/ours.err
