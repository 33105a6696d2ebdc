#!/bin/bash
# Round 2, GPU call 3 (1 GPU): validate fused GN backward / folded coefficients / PRO arrive fix, benches, ncu --set full.
mkdir -p gpurun_out; O=gpurun_out/r2_03; mkdir -p $O
echo "== tests"; timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_multi.py 2>&1 | tee $O/pytest_gpu_full.txt | tail -6
b() { tag=$1; shift; timeout 400 python bench.py --steps 20 --warmup 5 "$@" 2> $O/$tag.err | tee $O/$tag.json | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); a=d.get('alt') or {}
print('$tag', d['value'], d['ms_per_step'], d['dtype'], '| alt', a.get('dtype'), a.get('value'), a.get('ms_per_step'), '| loss', round(d['final_loss_acc'],3), 'launches/step', d['gpu_launches']/d['steps'], d['detail']['graph_nodes'])"; tail -2 $O/$tag.err; }
echo "== benches (tf32 headline + bf16 alt)"
b n1_b512
b n1_b64 --batch 64
b n1_b128 --batch 128
DLB_FUSED_GN_BWD=0 b nofusedgn_b64 --batch 64 --alt-dtype ""  --dtype bf16
DLB_FUSED_DGRAD=1 b fd_b512 --alt-dtype "" --dtype bf16
echo "== kernel bench (defaults)"; timeout 300 python tools/bench_kernels.py 2>&1 | tee $O/kernel_bench.txt | tail -12
echo "== gemm bench"; timeout 300 python tools/bench_gemm.py 2>&1 | tee $O/gemm_bench.txt | tail -30
echo "== eager launch list b=64 bf16"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3200 --csv --log-file $O/launches_b64.csv \
    python bench.py --batch 64 --steps 2 --warmup 5 --no-graphs --dtype bf16 --alt-dtype "" > $O/ncu_b64.log 2>&1
python tools/launch_summary.py $O/launches_b64.csv 2>&1 | tee $O/launches_b64.txt | head -20
echo "== ncu --set full: GN-prologue GEMM, 3x3 wgrad, fused GN backward (B=512 shapes)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_kernel|wgrad3x3_tc_kernel|gn_bwd_fused_kernel" -s 400 -c 12 \
    -o $O/prof_kernels python bench.py --batch 512 --steps 1 --warmup 5 --no-graphs --dtype bf16 --alt-dtype "" > $O/ncu_full.log 2>&1
ls -la $O/*.ncu-rep 2>/dev/null
