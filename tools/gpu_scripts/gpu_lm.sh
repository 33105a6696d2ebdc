#!/bin/bash
mkdir -p gpurun_out
echo "== tests"; timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -k "linear_cross or transformer or attention or layer_norm" 2>&1 | tail -4
echo "== transformer B=512 ours"; timeout 600 python bench.py --model transformer --batch 512 --steps 10 --warmup 3 2> gpurun_out/c_lm.err | tee gpurun_out/bench_ours_transformer_n1.json | cut -c1-260; tail -3 gpurun_out/c_lm.err
echo "== launches"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 1500 --csv --log-file gpurun_out/launches_lm.csv python bench.py --model transformer --batch 512 --steps 2 --warmup 3 --no-graphs > gpurun_out/ncu_lm.log 2>&1; wc -l gpurun_out/launches_lm.csv
