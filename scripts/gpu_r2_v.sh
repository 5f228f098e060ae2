#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu --timeout 300 -k "screened" -x > gpurun_out/v_p1.log 2>&1; tail -5 gpurun_out/v_p1.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sinkhorn_kernel -s 1 -c 1 -o gpurun_out/v_sk_screened python scripts/c4_once.py 1 > gpurun_out/v_ncu.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/v_ncu.log
