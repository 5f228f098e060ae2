#!/bin/bash
# C4 (float64-potential Sinkhorn): timing + one ncu --set full capture of the generic kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 200 python scripts/c4_once.py 5 2>&1 | tail -2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sinkhorn_kernel -s 1 -c 1 -o gpurun_out/t_sk_precise python scripts/c4_once.py 1 > gpurun_out/t_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/t_ncu.log
ls -la gpurun_out/*.ncu-rep
