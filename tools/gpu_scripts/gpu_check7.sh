#!/bin/bash
mkdir -p gpurun_out
echo "== gemm tests"; timeout 600 python -m pytest tests/test_gpu_gemm.py -x -q 2>&1 | tail -5
echo "== rest of gpu tests"; timeout 1200 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_gemm.py 2>&1 | tail -12
echo "== bench ours"; timeout 600 python bench.py --steps 10 --warmup 3 2> gpurun_out/bench_ours.err | tee gpurun_out/bench_ours.json | cut -c1-330; tail -3 gpurun_out/bench_ours.err
echo "== ncu launches (eager, steady state)"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 5000 -c 3000 --csv --log-file gpurun_out/launches4.csv python bench.py --steps 2 --warmup 3 --no-graphs > gpurun_out/ncu_bench.log 2>&1; tail -1 gpurun_out/ncu_bench.log | cut -c1-200; wc -l gpurun_out/launches4.csv
