#!/bin/bash
# Race / memory / synchronisation checking of the hand-written kernels with compute-sanitizer.
#   gpurun --timeout 1500 -- bash tools/gpu_scripts/sanitize.sh            # single-GPU kernels
#   gpurun --gpus 2 --timeout 1500 -- bash tools/gpu_scripts/sanitize.sh comm   # + the symmetric-memory collectives
# Summaries land in gpurun_out/sanitize_*.txt (copy the ones to keep into profiles/).
mkdir -p gpurun_out
SAN="compute-sanitizer --error-exitcode 9 --print-limit 20"
run() {   # name, tool, pytest selection...
  local name=$1 tool=$2; shift 2
  echo "== $tool: $name"
  timeout 900 $SAN --tool $tool --target-processes all python -m pytest "$@" -x -q -p no:cacheprovider \
      > gpurun_out/sanitize_${name}_${tool}.txt 2>&1
  echo "rc=$? $(grep -E 'ERROR SUMMARY|passed|failed' gpurun_out/sanitize_${name}_${tool}.txt | tail -2 | tr '\n' ' ')"
}
if [ "$1" != "comm-only" ]; then
# GroupNorm / pooling / optimizer / augment kernels (shared-memory reductions -> racecheck is the interesting tool)
run norm memcheck  tests/test_gpu_kernels.py -k "groupnorm or gn_ or pool or sgd or pack or augment"
run norm racecheck tests/test_gpu_kernels.py -k "groupnorm or gn_ or pool or sgd or pack or augment"
# tcgen05 GEMMs: memcheck + synccheck (mbarrier / bulk-async misuse); racecheck does not model the async proxy
run gemm memcheck  tests/test_gpu_gemm.py -k "plain_gemm or prologue or wgrad"
run gemm synccheck tests/test_gpu_gemm.py -k "plain_gemm or prologue or wgrad"
fi
if [ "$1" = "comm" ] || [ "$1" = "comm-only" ]; then
  run comm memcheck  tests/test_gpu_multi.py -k "allreduce or gather or barrier"
  run comm racecheck tests/test_gpu_multi.py -k "allreduce or gather or barrier"
fi
