#!/bin/bash
mkdir -p gpurun_out
echo "== tests"; timeout 1500 python -m pytest tests/test_gpu_kernels.py -x -q -k "linear_cross or transformer or attention or layer_norm or pack" 2>&1 | tail -4
echo "== transformer B=512 ours"; timeout 600 python bench.py --model transformer --batch 512 --steps 10 --warmup 3 2> gpurun_out/c_lm.err | tee gpurun_out/bench_ours_transformer_n1.json | cut -c1-200; tail -3 gpurun_out/c_lm.err
echo "== conv3 microbench"; timeout 300 python tools/bench_gemm.py 2>&1 | tail -7
