#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python scripts/ode_ab.py > gpurun_out/k_ode_ab.log 2>&1; cat gpurun_out/k_ode_ab.log
timeout 300 python scripts/ode_only.py > gpurun_out/k_ode_only.log 2>&1; tail -1 gpurun_out/k_ode_only.log
timeout 600 python -m pytest tests/test_gpu_round2.py -q -m gpu --timeout 300 -k "rk_stage or dopri5" > gpurun_out/k_p1.log 2>&1; tail -5 gpurun_out/k_p1.log
