#!/bin/bash
timeout 300 python - <<'PY'
import sys, torch, numpy as np
sys.path.insert(0, '.')
import cfm_b200
from cfm_b200 import _ffi
dev = torch.device('cuda:0')
n1 = 8192
for n0 in (148, 592, 1184, 2368, 4736, 8192):
    M = (torch.rand(n0, n1, device=dev) * 0.4 + 0.6).contiguous()
    cmax = M.max().reshape(1).contiguous()
    s = cfm_b200.OTPlanSampler('sinkhorn', reg=0.05, num_iter_max=100, stop_thr=0.0, warn=False, precision='fp32')
    ts = []
    for i in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); cp = s._solve_sinkhorn(M, cmax, n0, n1, 0.05, False); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    print(f"n0={n0:5d} rows/CTA={n0/148:6.1f}  {min(ts)*10:.2f} us/iter  variant={cp.status.cpu().tolist()[3]}")
PY
