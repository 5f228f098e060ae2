#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:ode_small --launch-skip 1 -c 1 -f -o gpurun_out/ode_small python scripts/ode_c1_once.py 2>&1 | tail -5
ls -la gpurun_out/ode_small.ncu-rep
