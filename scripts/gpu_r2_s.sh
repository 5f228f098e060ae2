#!/bin/bash
# stage-input kernel (templated on the stage): parity, then L2-policy x grid sweep on the C3 trajectory
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 300 -k "ode or dopri5 or trajectory or rk or stage" -x > gpurun_out/s_p1.log 2>&1; tail -4 gpurun_out/s_p1.log
for l2 in 0 3 7 4; do for g in 2 4 8; do
  echo "== CFM_RK_L2=$l2 CFM_RK_GRID=$g"
  CFM_RK_L2=$l2 CFM_RK_GRID=$g timeout 200 python scripts/ode_fuse_ab.py 00 2>&1 | grep "ms per" | tee -a gpurun_out/s_sweep.log
done; done
