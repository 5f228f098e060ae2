#!/bin/bash
# RK-stage fused MLP session: the new parity tests, the ODE tests, then the C3 A/B
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu --timeout 200 -k "stage_input_formed or rkstage" -x > gpurun_out/p_p0.log 2>&1; tail -15 gpurun_out/p_p0.log
timeout 900 python -m pytest tests -q -m gpu --timeout 300 -k "mlp or MLP or ode or dopri5 or trajectory or rk or fused or vector_field" -x > gpurun_out/p_p1.log 2>&1; tail -5 gpurun_out/p_p1.log
timeout 300 python scripts/ode_fuse_ab.py > gpurun_out/p_ab.log 2>&1; cat gpurun_out/p_ab.log | tail -12
