#!/bin/bash
mkdir -p gpurun_out
echo "== gemm tests"; timeout 600 python -m pytest tests/test_gpu_gemm.py -x -q 2>&1 | tail -5
echo "== gemm bench"; timeout 300 python tools/bench_gemm.py 2>&1 | tail -8
echo "== rest of gpu tests"; timeout 1200 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_gemm.py 2>&1 | tail -6
echo "== bench ours"; timeout 600 python bench.py --steps 10 --warmup 3 2> gpurun_out/bench_ours.err | tee gpurun_out/bench_ours.json | cut -c1-330; tail -3 gpurun_out/bench_ours.err
echo "== ncu full gemm_tc"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_kernel" -s 2 -c 2 -o gpurun_out/prof_gemm_tc python tools/bench_gemm.py quick > gpurun_out/ncu_gemm.log 2>&1; tail -2 gpurun_out/ncu_gemm.log
echo "== ncu launches (eager, steady state)"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 5000 -c 3000 --csv --log-file gpurun_out/launches5.csv python bench.py --steps 2 --warmup 3 --no-graphs > gpurun_out/ncu_bench.log 2>&1; wc -l gpurun_out/launches5.csv
