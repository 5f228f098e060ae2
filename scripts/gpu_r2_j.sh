#!/bin/bash
# round-2 session j: overlapped stage inputs (side stream): ODE parity tests + A/B timing
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
echo "== ode / mlp tests" | tee gpurun_out/j_p1.log
timeout 900 python -m pytest tests -q -m gpu --timeout 400 -k "dopri5 or mlp or fused_small or smoke" >> gpurun_out/j_p1.log 2>&1
echo "rc=$?" >> gpurun_out/j_p1.log; tail -8 gpurun_out/j_p1.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/j_smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 python scripts/ode_ab.py > gpurun_out/j_ode_ab.log 2>&1; cat gpurun_out/j_ode_ab.log
