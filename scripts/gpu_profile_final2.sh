#!/bin/bash
mkdir -p gpurun_out
echo "== full: gemm_tc sqdist"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 3 -c 1 -o gpurun_out/prof_sqdist_tc python bench.py --steps 1 --warmup 3 --no-ode --no-cpu-baseline > gpurun_out/ncu_full2.log 2>&1; echo "rc=$?"
echo "== full: gemm_tc mlp"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 40 -c 4 -o gpurun_out/prof_mlp_tc python scripts/ode_only.py --eager > gpurun_out/ncu_full3.log 2>&1; echo "rc=$?"
ls -la gpurun_out/*.ncu-rep
