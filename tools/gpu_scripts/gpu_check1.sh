#!/bin/bash
# First GPU contact: smoke, GPU tests, both bench arms, launch list.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25
echo "== bench ours"; timeout 600 python bench.py --steps 10 --warmup 3 2> gpurun_out/bench_ours.err | tee gpurun_out/bench_ours.json; tail -5 gpurun_out/bench_ours.err
echo "== bench ours nograph"; timeout 600 python bench.py --steps 10 --warmup 3 --no-graphs 2> gpurun_out/bench_ours_ng.err | tee gpurun_out/bench_ours_nograph.json; tail -3 gpurun_out/bench_ours_ng.err
echo "== bench reference"; timeout 900 python bench.py --impl reference --steps 10 --warmup 3 2> gpurun_out/bench_ref.err | tee gpurun_out/bench_ref.json; tail -5 gpurun_out/bench_ref.err
echo "== ncu launches"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-graphs > gpurun_out/ncu_bench.log 2>&1; tail -2 gpurun_out/ncu_bench.log; wc -l gpurun_out/launches.csv
