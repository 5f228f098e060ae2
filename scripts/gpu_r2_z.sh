#!/bin/bash
# sinkhorn_v2 combine variants: 4-wide loop (CFM_SK_DBG=1, default so far), all partials in one round trip (0), 8-wide loop (4)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for dbg in 1 0 4 1 0; do
  CFM_SK_DBG=$dbg timeout 300 python bench.py --steps 20 --warmup 3 --no-ode --no-cpu-baseline --no-extra > gpurun_out/z_$dbg.json 2>/dev/null
  python -c "
import json; d=json.loads(open('gpurun_out/z_$dbg.json').read().strip().splitlines()[-1])
print('CFM_SK_DBG=$dbg', 'value', round(d['value'],2), 'ms', round(d['ms_per_step'],3), 'solve', round(d['roofline']['stage_ms']['solve'],4), 'parity', d.get('parity'))"
done
