#!/bin/bash
# MLP-focused session: parity tests that touch the MLP / ODE paths, then the fused-kernel timeline
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 300 -k "mlp or MLP or ode or dopri5 or trajectory or rk or fused or vector_field" -x > gpurun_out/m_p1.log 2>&1; tail -8 gpurun_out/m_p1.log
PROBES="${PROBES:-0}" bash scripts/gpu_r2_tl.sh 2>&1 | grep -v "^sqdist\|^cost stage"
timeout 300 python scripts/ode_only.py > gpurun_out/m_ode_only.log 2>&1; tail -2 gpurun_out/m_ode_only.log
