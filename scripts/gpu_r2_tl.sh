#!/bin/bash
# fused-MLP / cost-GEMM timelines only
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for probe in ${PROBES:-0}; do
  echo "== CFM_MLP_PROBE=$probe"
  CFM_MLP_PROBE=$probe timeout 300 python scripts/mlp_timeline.py 2>&1 | tee gpurun_out/tl_probe$probe.log
done
