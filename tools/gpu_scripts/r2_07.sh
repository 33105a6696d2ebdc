#!/bin/bash
# Round 2, GPU call 7 (1 GPU): bulk-copy (TMA unit) fused GroupNorm backward -- tests, flavour microbench, step A/B;
# full GPU suite on the final tree; ncu --set full of the block-1 kernels (tf32), kept small enough to travel back.
mkdir -p gpurun_out; O=gpurun_out/r2_07; mkdir -p $O
echo "== tests (one process per file, no -x)"
for f in tests/test_gpu_kernels.py tests/test_gpu_dgrad_gn.py tests/test_gpu_gemm.py tests/test_gpu_gemm_tf32.py tests/test_gpu_gradpath.py; do
  n=$(basename $f .py); timeout 420 python -m pytest $f -m gpu -q 2>&1 | tee $O/pytest_$n.txt | grep -E "FAILED|passed|failed|error" | head -12
done
echo "== fused GN backward: register vs bulk-copy flavour"
timeout 300 python tools/bench_kernels.py gnbwd 2>&1 | tee $O/gn_bwd_flavours.txt | tail -24
b() { tag=$1; shift; timeout 400 python bench.py --steps 20 --warmup 5 "$@" 2> $O/$tag.err | tee $O/$tag.json | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); a=d.get('alt') or {}
print('$tag', d['value'], d['ms_per_step'], d['dtype'], '| alt', a.get('dtype'), a.get('value'), a.get('ms_per_step'), '| loss', round(d['final_loss_acc'],3), 'launches/step', d['gpu_launches']/d['steps'], d['detail']['graph_nodes'])"; tail -2 $O/$tag.err; }
echo "== benches (tf32 headline + bf16 alt)"
b n1_b512
DLB_GN_BWD_BULK=0 b nobulk_b512
b n1_b64 --batch 64
DLB_GN_BWD_BULK=0 b nobulk_b64 --batch 64
echo "== ncu --set full: block-1 shapes, tf32 (one dense layer, forward + backward)"
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -c 12 \
    -k regex:"gemm_tc_kernel|dgrad_gn_kernel|wgrad_tc_kernel|wgrad3x3_tc_kernel|gn_bwd_bulk_kernel|gn_bwd_fused_kernel|gn_fwd_apply_kernel" \
    -o $O/prof_block1_tf32 python tools/ncu_targets.py --dtype tf32 --batch 256 --layers 1 > $O/ncu_full_tf32.log 2>&1
tail -2 $O/ncu_full_tf32.log
python tools/ncu_summary.py $O/prof_block1_tf32.ncu-rep > $O/ncu_block1_tf32_summary.txt 2>&1; grep -c "^kernel" $O/ncu_block1_tf32_summary.txt
ls -la $O/*.ncu-rep; s=$(stat -c %s $O/prof_block1_tf32.ncu-rep 2>/dev/null || echo 0); if [ "$s" -gt 30000000 ]; then rm -f $O/prof_block1_tf32.ncu-rep; echo "rep too large, removed"; fi
du -sh gpurun_out
