#!/bin/bash
mkdir -p gpurun_out
timeout 300 python - <<'PY'
import sys, torch, numpy as np
sys.path.insert(0, '.')
import cfm_b200
dev = torch.device('cuda:0')
x0, x1 = torch.randn(8192, 784, device=dev), torch.randn(8192, 784, device=dev)
for algo in (0, 2, 1):
    s = cfm_b200.OTPlanSampler('sinkhorn', reg=0.05, normalize_cost=True, cost_algo=algo, warn=False)
    ts = []
    for i in range(8):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); s._cost(x0, x1, dev); b.record(); torch.cuda.synchronize(); ts.append(round(a.elapsed_time(b), 3))
    print('algo', algo, ts)
s = cfm_b200.OTPlanSampler('sinkhorn', reg=0.05, normalize_cost=True, num_iter_max=100, stop_thr=0.0, warn=False)
for rep in range(2):
    s.stage_events = []
    for i in range(6): s.sample_plan(x0, x1)
    torch.cuda.synchronize()
    d = {}
    for n, a, b in s.stage_events: d.setdefault(n, []).append(round(a.elapsed_time(b), 3))
    print(d)
PY
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['roofline']['stage_ms'], d['ode']['value'], d['ode']['ms_per_trajectory'])"
