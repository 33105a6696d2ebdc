#!/bin/bash
# First GPU call of round 2: validate the default path, then the experimental fused dgrad+GN-backward kernel
# (csrc/dgrad_gn.cu, OFF by default), its micro-benchmark and an A/B of the headline bench with the flag on.
#   gpurun --timeout 1500 -- bash tools/gpu_scripts/round2_first.sh
mkdir -p gpurun_out
echo "== default gemm tests";      timeout 300 python -m pytest tests/test_gpu_gemm.py -x -q 2>&1 | tail -2
echo "== experimental tests (own process: a protocol bug traps the context)"
DLB_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_experimental.py -x -q 2>&1 | tail -15
echo "== microbench chain vs fused"; timeout 300 python tools/bench_dgrad_gn.py 2>&1 | tail -10 | tee gpurun_out/bench_dgrad_gn.txt
echo "== bench default";            timeout 300 python bench.py --steps 10 --warmup 5 2> gpurun_out/b_default.err | tee gpurun_out/b_default.json | cut -c1-200
echo "== bench DLB_FUSED_DGRAD=1";  DLB_FUSED_DGRAD=1 timeout 300 python bench.py --steps 10 --warmup 5 2> gpurun_out/b_fused.err | tee gpurun_out/b_fused.json | cut -c1-200
tail -3 gpurun_out/b_fused.err
