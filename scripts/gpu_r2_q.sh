#!/bin/bash
# RK-stage fused MLP, prefetch distance sweep
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m pytest tests -q -m gpu --timeout 200 -k "stage_input_formed or rkstage" -x > gpurun_out/q_p0.log 2>&1; tail -3 gpurun_out/q_p0.log
for pf in 0 2 4; do
  CFM_RK_PF=$pf timeout 200 python scripts/rk_stage_timeline.py 2>&1 | tee gpurun_out/q_tl_pf$pf.log
done
CFM_RK_PF=2 timeout 200 python scripts/ode_fuse_ab.py 2>&1 | tee gpurun_out/q_ab.log | head -3
