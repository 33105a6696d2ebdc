#!/bin/bash
# Round 2, GPU call 1 (1 GPU): default suite, experimental fused dgrad+GN kernel, both bench arms, graph node counts.
#   gpurun --timeout 1500 -- bash tools/gpu_scripts/r2_01_validate.sh
mkdir -p gpurun_out; O=gpurun_out/r2_01; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu.txt
echo "== tf32 gemm tests"; timeout 600 python -m pytest tests/test_gpu_gemm_tf32.py -q 2>&1 | tail -15 | tee $O/pytest_tf32.txt
echo "== gpu tests";       timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_gemm_tf32.py 2>&1 | tail -8 | tee $O/pytest_gpu.txt
echo "== experimental";    timeout 300 python -m pytest tests/test_gpu_dgrad_gn.py -x -q 2>&1 | tail -8 | tee $O/pytest_exp.txt
echo "== microbench";      timeout 300 python tools/bench_dgrad_gn.py 2>&1 | tail -10 | tee $O/bench_dgrad_gn.txt
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 20 --warmup 5 2> $O/ref.err | tee $O/ref.json | cut -c1-300
echo "== bench ours";      timeout 300 python bench.py --steps 20 --warmup 5 2> $O/ours.err | tee $O/ours.json | cut -c1-300
echo "== bench ours tf32";  timeout 300 python bench.py --steps 20 --warmup 5 --dtype tf32 2> $O/ours_tf32.err | tee $O/ours_tf32.json | cut -c1-300
echo "== bench ours FUSED_DGRAD"; DLB_FUSED_DGRAD=1 timeout 300 python bench.py --steps 20 --warmup 5 2> $O/ours_fd.err | tee $O/ours_fd.json | cut -c1-300
for b in 512 64; do
  echo "== graph nodes, batch $b"
  DLB_GRAPH_DUMP=$O/graph timeout 300 python bench.py --batch $b --steps 10 --warmup 5 2> $O/step_b$b.err | tee $O/step_b$b.json | cut -c1-200
  f=$O/graph.b$b.rank0.dot
  [ -f $f ] && python tools/graph_nodes.py $f --top 30 | tee $O/graph_b$b.txt | head -14
  rm -f $O/graph.b*.dot
done
echo "== eager launch list, batch 64"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/launches_b64.csv \
    python bench.py --batch 64 --steps 2 --warmup 5 --no-graphs > $O/ncu_b64.log 2>&1
python tools/launch_summary.py $O/launches_b64.csv 2>&1 | tee $O/launches_b64.txt | head -24
tail -3 $O/*.err | tail -30
