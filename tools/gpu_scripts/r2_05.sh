#!/bin/bash
# Round 2, GPU call 5 (1 GPU, short): validate stem rewrite + staged GN backward before the 8-GPU call.
mkdir -p gpurun_out; O=gpurun_out/r2_05; mkdir -p $O
echo "== tests"; timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_gradpath.py tests/test_gpu_gemm_tf32.py -q -x 2>&1 | tee $O/pytest.txt | tail -5
echo "== experimental"; timeout 300 python -m pytest tests/test_gpu_dgrad_gn.py -q 2>&1 | tail -4 | tee $O/pytest_exp.txt
b() { tag=$1; shift; timeout 400 python bench.py --steps 20 --warmup 5 "$@" 2> $O/$tag.err | tee $O/$tag.json | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); a=d.get('alt') or {}
print('$tag', d['value'], d['ms_per_step'], d['dtype'], '| alt', a.get('dtype'), a.get('value'), a.get('ms_per_step'), '| loss', round(d['final_loss_acc'],3), d['detail']['graph_nodes'])"; tail -2 $O/$tag.err; }
b n1_b512
DLB_GN_BWD_STAGE=0 b nostage_b512
b n1_b64 --batch 64
echo "== smoke"; timeout 300 python __graft_entry__.py 2>&1 | tail -2
