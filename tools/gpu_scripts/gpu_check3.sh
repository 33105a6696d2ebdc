#!/bin/bash
mkdir -p gpurun_out
echo "== kernel microbench"; timeout 600 python tools/bench_kernels.py 2>&1 | tail -60
echo "== ncu full on the GN kernels"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"nc_reduce2|gn_fwd_apply|gn_bwd_apply" -s 12 -c 8 -o gpurun_out/prof_gn python tools/bench_kernels.py quick > gpurun_out/ncu_gn.log 2>&1; tail -3 gpurun_out/ncu_gn.log; ls -la gpurun_out/*.ncu-rep
