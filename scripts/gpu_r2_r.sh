#!/bin/bash
# RK-stage fused MLP: load-mode sweep; stage-input kernel: L2 policy sweep
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for ld in 1 2; do for pf in 0 2; do
  CFM_RK_LD=$ld CFM_RK_PF=$pf timeout 200 python scripts/rk_stage_timeline.py 2>&1 | grep -v "^stage [2345]:" | tee gpurun_out/r_tl_ld${ld}_pf$pf.log
done; done
for l2 in 0 1 2 3 7; do
  CFM_RK_L2=$l2 timeout 200 python scripts/ode_fuse_ab.py 00 2>&1 | tee gpurun_out/r_ab_l2_$l2.log
done
CFM_RK_LD=1 CFM_RK_PF=0 timeout 200 python scripts/ode_fuse_ab.py 1 2>&1 | tee gpurun_out/r_ab_fuse.log
