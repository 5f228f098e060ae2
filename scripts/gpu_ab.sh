#!/bin/bash
for rep in 1 2; do for dbg in 0 3 1 2; do
CFM_SK_DBG=$dbg timeout 300 python - <<PY
import sys, torch
sys.path.insert(0, '.')
import cfm_b200
dev = torch.device('cuda:0')
n0 = n1 = 8192
torch.manual_seed(0)
M = (torch.rand(n0, n1, device=dev) * 0.4 + 0.6).contiguous(); cmax = M.max().reshape(1).contiguous()
s = cfm_b200.OTPlanSampler('sinkhorn', reg=0.05, num_iter_max=100, stop_thr=0.0, warn=False, precision='fp32')
ts = []
for i in range(8):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); cp = s._solve_sinkhorn(M, cmax, n0, n1, 0.05, False); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
print('dbg $dbg', ' '.join(f'{t*10:.2f}' for t in ts[2:]))
PY
done; done
