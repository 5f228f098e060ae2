#!/bin/bash
# Final evidence pass: launch list of the bench command + one full capture per dominant kernel.
mkdir -p gpurun_out
rm -f gpurun_out/prof_*.ncu-rep
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"
echo "== launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "rc=$?"
echo "== full: sinkhorn_v2"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:sinkhorn_v2 -s 3 -c 1 -o gpurun_out/prof_sinkhorn_v2 python bench.py --steps 1 --warmup 3 --no-ode --no-cpu-baseline > gpurun_out/ncu_full1.log 2>&1; echo "rc=$?"
echo "== full: gemm_tc sqdist"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:SqDistTc -s 3 -c 1 -o gpurun_out/prof_sqdist_tc python bench.py --steps 1 --warmup 3 --no-ode --no-cpu-baseline > gpurun_out/ncu_full2.log 2>&1; echo "rc=$?"
echo "== full: gemm_tc mlp (layer 0)"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:MlpTc -s 40 -c 1 -o gpurun_out/prof_mlp_tc python scripts/ode_only.py --eager > gpurun_out/ncu_full3.log 2>&1; echo "rc=$?"
echo "== full: draw"; timeout 900 ncu --set full --clock-control none -k regex:draw_uniform -s 3 -c 1 -o gpurun_out/prof_draw python bench.py --steps 1 --warmup 3 --no-ode --no-cpu-baseline > gpurun_out/ncu_full4.log 2>&1; echo "rc=$?"
ls -la gpurun_out/*.ncu-rep
