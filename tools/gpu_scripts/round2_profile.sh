#!/bin/bash
# Where does a step go, at the headline batch and at the small per-rank batches of N = 4/8?  One GPU, one call:
#   gpurun --timeout 1500 -- bash tools/gpu_scripts/round2_profile.sh
# Produces in gpurun_out/: step_b*.json (graph-mode step time per local batch), graph_b*.txt (graph node counts,
# kernel histogram, critical-path length), launches_b*.csv + launches_b*.txt (ncu kernel list of an eager step).
mkdir -p gpurun_out
for b in 512 256 128 64; do
  echo "== graph-mode step, batch $b"
  DLB_GRAPH_DUMP=gpurun_out/graph timeout 300 python bench.py --batch $b --steps 10 --warmup 5 2> gpurun_out/step_b$b.err \
      | tee gpurun_out/step_b$b.json | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print(d['ms_per_step'], 'ms/step', d['value'], d['unit'], 'launches/step', d['gpu_launches']/d['steps'])"
  f=gpurun_out/graph.b$b.rank0.dot
  [ -f $f ] && python tools/graph_nodes.py $f --top 30 | tee gpurun_out/graph_b$b.txt | head -12
  rm -f gpurun_out/graph.b*.dot        # the dot files are tens of MB
done
for b in 512 64; do
  echo "== eager launch list, batch $b"
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_b$b.csv \
      python bench.py --batch $b --steps 2 --warmup 4 --no-graphs > gpurun_out/ncu_b$b.log 2>&1
  python tools/launch_summary.py gpurun_out/launches_b$b.csv 2>&1 | tee gpurun_out/launches_b$b.txt | head -14
done
