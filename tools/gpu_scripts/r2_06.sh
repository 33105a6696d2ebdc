#!/bin/bash
# Round 2, GPU call 6 (1 GPU): tf32 fused dgrad+GN backward, any-batch 3x3 wgrad, tcgen05 classifier heads, stem / max-pool tests;
# benches with A/B switches; ncu --set full of the block-1 kernels; launch list of the tf32 step.
mkdir -p gpurun_out; O=gpurun_out/r2_06; mkdir -p $O
echo "== tests (one process per file, no -x)"
for f in tests/test_gpu_dgrad_gn.py tests/test_gpu_gemm.py tests/test_gpu_gemm_tf32.py tests/test_gpu_kernels.py tests/test_gpu_gradpath.py; do
  n=$(basename $f .py); timeout 420 python -m pytest $f -m gpu -q 2>&1 | tee $O/pytest_$n.txt | grep -E "FAILED|passed|failed|error" | head -12
done
b() { tag=$1; shift; timeout 400 python bench.py --steps 20 --warmup 5 "$@" 2> $O/$tag.err | tee $O/$tag.json | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); a=d.get('alt') or {}
print('$tag', d['value'], d['ms_per_step'], d['dtype'], '| alt', a.get('dtype'), a.get('value'), a.get('ms_per_step'), '| loss', round(d['final_loss_acc'],3), 'launches/step', d['gpu_launches']/d['steps'], d['detail']['graph_nodes'])"; tail -2 $O/$tag.err; }
echo "== benches (tf32 headline + bf16 alt)"
b n1_b512
DLB_FUSED_DGRAD_TF32=0 b nofd32_b512 --alt-dtype ""
b n1_b64 --batch 64
DLB_FUSED_DGRAD_TF32=0 b nofd32_b64 --batch 64 --alt-dtype ""
b n1_b67 --batch 67 --alt-dtype ""
echo "== ncu --set full: block-1 shapes, tf32"
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -c 40 -o $O/prof_block1_tf32 \
    python tools/ncu_targets.py --dtype tf32 --batch 256 --layers 3 > $O/ncu_full_tf32.log 2>&1
tail -2 $O/ncu_full_tf32.log
python tools/ncu_summary.py $O/prof_block1_tf32.ncu-rep > $O/ncu_block1_tf32_summary.txt 2>&1; grep -c "^kernel" $O/ncu_block1_tf32_summary.txt
echo "== eager launch list b=512 tf32"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2800 --csv --log-file $O/launches_b512_tf32.csv \
    python bench.py --batch 512 --steps 2 --warmup 5 --no-graphs --dtype tf32 --alt-dtype "" > $O/ncu_b512.log 2>&1
python tools/launch_summary.py $O/launches_b512_tf32.csv 2>&1 | tee $O/launches_b512_tf32.txt | head -18
echo "== smoke"; timeout 300 python __graft_entry__.py 2>&1 | tail -2
