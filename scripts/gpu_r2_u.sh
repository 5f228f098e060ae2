#!/bin/bash
# float64-potential Sinkhorn (streamed seeded screening): parity then C4 timing against the other variants
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 300 -k "sinkhorn or screened or c4 or draw" -x > gpurun_out/u_p1.log 2>&1; tail -6 gpurun_out/u_p1.log
timeout 200 python scripts/c4_once.py 5 2>&1 | tail -1
CFM_SK_STREAM=0 timeout 200 python scripts/c4_once.py 5 2>&1 | tail -1
CFM_SK_SCREEN=0 timeout 200 python scripts/c4_once.py 5 2>&1 | tail -1
timeout 200 python scripts/c4_timeline.py 2>&1 | tail -2
