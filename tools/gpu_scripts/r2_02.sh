#!/bin/bash
# Round 2, GPU call 2 (1 GPU): tf32 MN-major fix, new kernels (3x3 wgrad, sinks, CE, embedding, SE), GN sweep, A/B benches.
mkdir -p gpurun_out; O=gpurun_out/r2_02; mkdir -p $O
echo "== tf32 gemm tests"; timeout 600 python -m pytest tests/test_gpu_gemm_tf32.py -q 2>&1 | tee $O/pytest_tf32_full.txt | tail -12
echo "== wgrad3x3 + gemm tests"; timeout 600 python -m pytest tests/test_gpu_gemm.py -q 2>&1 | tee $O/pytest_gemm_full.txt | tail -8
echo "== grad path / kernels"; timeout 900 python -m pytest tests/test_gpu_gradpath.py tests/test_gpu_kernels.py -q 2>&1 | tee $O/pytest_kernels_full.txt | tail -12
echo "== GN sweep"; timeout 600 python tools/bench_kernels.py sweep 2>&1 | tee $O/gn_sweep.txt | tail -5
b() { tag=$1; shift; timeout 300 python bench.py --steps 20 --warmup 5 "$@" 2> $O/$tag.err | tee $O/$tag.json | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('$tag', d['value'], d['ms_per_step'], d['dtype'], 'loss', round(d['final_loss_acc'],3), 'launches/step', d['gpu_launches']/d['steps'], d['detail']['graph_nodes'])"; tail -2 $O/$tag.err; }
echo "== benches"
b ours_b512
b ours_b64 --batch 64
DLB_GRAD_SINKS=0 b nosink_b512
DLB_GRAD_SINKS=0 b nosink_b64 --batch 64
DLB_TC_WGRAD3=0 b nowg3_b512
DLB_TC_WGRAD3=0 b nowg3_b64 --batch 64
b tf32_b512 --dtype tf32
b tf32_b64 --dtype tf32 --batch 64
DLB_FUSED_DGRAD=1 b fd_b64 --batch 64
echo "== eager launch list, batch 512 and 64 (new default path)"
for bb in 512 64; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3600 --csv --log-file $O/launches_b$bb.csv \
    python bench.py --batch $bb --steps 2 --warmup 5 --no-graphs > $O/ncu_b$bb.log 2>&1
python tools/launch_summary.py $O/launches_b$bb.csv 2>&1 | tee $O/launches_b$bb.txt | head -22
done
