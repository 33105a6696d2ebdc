#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
echo "== bench ours"; timeout 600 python bench.py --steps 10 --warmup 3 2> gpurun_out/bench_ours.err | tee gpurun_out/bench_ours.json; tail -5 gpurun_out/bench_ours.err
echo "== ncu launches (eager, steady state)"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 6000 -c 3000 --csv --log-file gpurun_out/launches2.csv python bench.py --steps 2 --warmup 3 --no-graphs > gpurun_out/ncu_bench.log 2>&1; tail -2 gpurun_out/ncu_bench.log | cut -c1-300; wc -l gpurun_out/launches2.csv
