#!/bin/bash
# final single-GPU validation
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== pytest -m gpu"; timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== oversubscribed GPU (-gpu 0,0 over gloo, like the reference's -gpu 0,0,0,1)"; cd /tmp && rm -rf ov && mkdir ov && cd ov && timeout 600 python $GRAFT_REPO_ROOT/dbs.py -d false -ws 2 -b 64 -m mnistnet -ds mnist -e 2 -gpu 0,0 --synthetic true --train_samples 1024 --test_samples 256 --master_port 29655 2>&1 | grep -E "Total Time|accuracy|Error|error" | tail -4; cd $GRAFT_REPO_ROOT
echo "== checkpoint + resume"; cd /tmp && rm -rf ck && mkdir ck && cd ck && timeout 600 python $GRAFT_REPO_ROOT/dbs.py -d false -ws 1 -b 64 -m resnet18 -ds cifar10 -e 2 -gpu 0 --synthetic true --train_samples 512 --test_samples 128 --checkpoint_dir ./ckpt --master_port 29656 2>&1 | grep -E "val_loss|Error" | tail -2; timeout 600 python $GRAFT_REPO_ROOT/dbs.py -d false -ws 1 -b 64 -m resnet18 -ds cifar10 -e 3 -gpu 0 --synthetic true --train_samples 512 --test_samples 128 --checkpoint_dir ./ckpt --resume true --master_port 29657 2>&1 | grep -E "resumed|val_loss|Error" | tail -3; cd $GRAFT_REPO_ROOT
echo "== bench ours N=1"; timeout 600 python bench.py --steps 20 --warmup 5 2> gpurun_out/f1.err | tee gpurun_out/final_bench_ours_n1.json | cut -c1-250; tail -2 gpurun_out/f1.err
echo "== bench ref N=1"; timeout 900 python bench.py --impl reference --steps 20 --warmup 5 2> gpurun_out/f1r.err | tee gpurun_out/final_bench_ref_n1.json | cut -c1-250
