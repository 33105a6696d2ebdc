#!/bin/bash
mkdir -p gpurun_out
echo "== resnet50 B=1024 ours"; timeout 600 python bench.py --model resnet50 --batch 1024 --steps 10 --warmup 3 2> gpurun_out/c_r50.err | tee gpurun_out/bench_ours_resnet50_n1.json | cut -c1-260; tail -2 gpurun_out/c_r50.err
echo "== resnet50 B=1024 ref"; timeout 900 python bench.py --impl reference --model resnet50 --batch 1024 --steps 6 --warmup 3 2> gpurun_out/c_r50r.err | tee gpurun_out/bench_ref_resnet50_n1.json | cut -c1-260; tail -2 gpurun_out/c_r50r.err
echo "== transformer B=512 ours"; timeout 600 python bench.py --model transformer --batch 512 --steps 10 --warmup 3 2> gpurun_out/c_lm.err | tee gpurun_out/bench_ours_transformer_n1.json | cut -c1-260; tail -3 gpurun_out/c_lm.err
echo "== transformer B=512 ref"; timeout 900 python bench.py --impl reference --model transformer --batch 512 --steps 10 --warmup 3 2> gpurun_out/c_lmr.err | tee gpurun_out/bench_ref_transformer_n1.json | cut -c1-260; tail -3 gpurun_out/c_lmr.err
