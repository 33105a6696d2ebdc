#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
echo "== kernel microbench"; timeout 600 python tools/bench_kernels.py 2>&1 | grep -v param_grad | tail -50
echo "== bench ours"; timeout 600 python bench.py --steps 10 --warmup 3 2> gpurun_out/bench_ours.err | tee gpurun_out/bench_ours.json | cut -c1-400; tail -5 gpurun_out/bench_ours.err
echo "== ncu launches (eager, steady state)"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 5000 -c 3000 --csv --log-file gpurun_out/launches3.csv python bench.py --steps 2 --warmup 3 --no-graphs > gpurun_out/ncu_bench.log 2>&1; tail -1 gpurun_out/ncu_bench.log | cut -c1-200; wc -l gpurun_out/launches3.csv
