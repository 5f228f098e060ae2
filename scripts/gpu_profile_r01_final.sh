#!/bin/bash
# final round-1 evidence: launch list of the bench command, full captures of the two kernels added late in the round
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:assign_fast --launch-skip 2 -c 1 -f -o gpurun_out/assign_fast python scripts/c1_coupling_once.py 2>&1 | tail -3
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_final.csv
