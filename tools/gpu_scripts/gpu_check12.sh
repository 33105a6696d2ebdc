#!/bin/bash
mkdir -p gpurun_out
echo "== gemm tests"; timeout 600 python -m pytest tests/test_gpu_gemm.py -x -q 2>&1 | tail -6
echo "== rest of gpu tests"; timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_gemm.py 2>&1 | tail -6
echo "== bench ours"; timeout 600 python bench.py --steps 10 --warmup 3 2> gpurun_out/bench_ours.err | tee gpurun_out/bench_ours.json | cut -c1-330; tail -3 gpurun_out/bench_ours.err
echo "== bench b128"; timeout 600 python bench.py --steps 20 --warmup 3 --batch 128 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ms/step', d['ms_per_step'], 'launches/step', d['gpu_launches']/d['steps'])"
echo "== ncu fused kernels"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_kernel|wgrad_tc_kernel" -s 40 -c 14 -o gpurun_out/prof_gemm_fused python tools/bench_gemm.py quick > gpurun_out/ncu_gemm.log 2>&1; ls -la gpurun_out/prof_gemm_fused.ncu-rep
