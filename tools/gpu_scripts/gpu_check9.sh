#!/bin/bash
mkdir -p gpurun_out
echo "== gemm tests"; timeout 600 python -m pytest tests/test_gpu_gemm.py -x -q 2>&1 | tail -4
echo "== gemm bench"; timeout 300 python tools/bench_gemm.py 2>&1 | tail -5
echo "== bench ours"; timeout 600 python bench.py --steps 10 --warmup 3 2> gpurun_out/bench_ours.err | tee gpurun_out/bench_ours.json | cut -c1-330; tail -3 gpurun_out/bench_ours.err
echo "== ncu full fused gemm_tc"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_kernel|wgrad_tc" -s 60 -c 12 -o gpurun_out/prof_gemm_fused python tools/bench_gemm.py quick > gpurun_out/ncu_gemm.log 2>&1; tail -2 gpurun_out/ncu_gemm.log
