#!/bin/bash
mkdir -p gpurun_out
echo "== gemm tests"; timeout 600 python -m pytest tests/test_gpu_gemm.py -x -q 2>&1 | tail -4
echo "== gemm bench"; timeout 300 python tools/bench_gemm.py 2>&1 | tail -4
echo "== GN kernel microbench"; timeout 600 python tools/bench_kernels.py 2>&1 | grep -E "C=  128|C=  256|C= 1024" | grep -v param_grad | head -30
echo "== rest of gpu tests"; timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_gemm.py 2>&1 | tail -8
echo "== bench ours"; timeout 600 python bench.py --steps 10 --warmup 3 2> gpurun_out/bench_ours.err | tee gpurun_out/bench_ours.json | cut -c1-330; tail -3 gpurun_out/bench_ours.err
