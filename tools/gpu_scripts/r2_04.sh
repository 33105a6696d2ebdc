#!/bin/bash
# Round 2, GPU call 4 (1 GPU): TMA-staged GN coefficients, GN atomics fix, halo 3x3, stem conv -> tests, benches, launch lists.
mkdir -p gpurun_out; O=gpurun_out/r2_04; mkdir -p $O
echo "== tests"; timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_multi.py 2>&1 | tee $O/pytest_gpu_full.txt | tail -8
b() { tag=$1; shift; timeout 400 python bench.py --steps 20 --warmup 5 "$@" 2> $O/$tag.err | tee $O/$tag.json | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); a=d.get('alt') or {}
print('$tag', d['value'], d['ms_per_step'], d['dtype'], '| alt', a.get('dtype'), a.get('value'), a.get('ms_per_step'), '| loss', round(d['final_loss_acc'],3), 'launches/step', d['gpu_launches']/d['steps'], d['detail']['graph_nodes'])"; tail -2 $O/$tag.err; }
echo "== benches (tf32 headline + bf16 alt)"
b n1_b512
b n1_b64 --batch 64
DLB_CONV3_HALO=0 b nohalo_b512 --dtype bf16 --alt-dtype ""
DLB_PRO_TMA=0 b noprotma_b64 --dtype bf16 --alt-dtype "" --batch 64
DLB_FUSED_GN_BWD=0 b nofusedgn_b64 --dtype bf16 --alt-dtype "" --batch 64
DLB_FUSED_DGRAD=1 b fd_b64 --dtype bf16 --alt-dtype "" --batch 64
echo "== gemm bench"; timeout 300 python tools/bench_gemm.py 2>&1 | tee $O/gemm_bench.txt | tail -16
for dt in bf16 tf32; do
echo "== eager launch list b=64 $dt"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2800 --csv --log-file $O/launches_b64_$dt.csv \
    python bench.py --batch 64 --steps 2 --warmup 5 --no-graphs --dtype $dt --alt-dtype "" > $O/ncu_b64_$dt.log 2>&1
python tools/launch_summary.py $O/launches_b64_$dt.csv 2>&1 | tee $O/launches_b64_$dt.txt | head -16
done
echo "== eager launch list b=512 bf16"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2800 --csv --log-file $O/launches_b512_bf16.csv \
    python bench.py --batch 512 --steps 2 --warmup 5 --no-graphs --dtype bf16 --alt-dtype "" > $O/ncu_b512.log 2>&1
python tools/launch_summary.py $O/launches_b512_bf16.csv 2>&1 | tee $O/launches_b512_bf16.txt | head -16
