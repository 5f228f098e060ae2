#!/bin/bash
timeout 600 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider -k "fused_flow" -x 2>&1 | grep -v "^$" | tail -40 | cut -c1-260
timeout 300 python - <<'PY'
import sys, time, torch, numpy as np
sys.path.insert(0, '.')
import cfm_b200
dev = torch.device('cuda:0')
for n, d in ((128, 8), (256, 2), (1024, 32), (2048, 64)):
    g = torch.Generator().manual_seed(n)
    a, b = torch.randn(n, d, generator=g).to(dev), torch.randn(n, d, generator=g).to(dev)
    s = cfm_b200.OTPlanSampler('exact', warn=False)
    for _ in range(2): s.sample_plan(a, b)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    reps = 20 if n <= 1024 else 3
    for _ in range(reps): s.sample_plan(a, b)
    torch.cuda.synchronize(); gpu = (time.perf_counter() - t0) / reps
    cp = s._couple(a, b, dev); st = cp.status.cpu().tolist()
    print(f"n={n} d={d} gpu {gpu*1e3:.3f} ms/coupling  augmentations={st[1]} dijkstra_steps={st[2]} us/step={gpu*1e6/max(1,st[2]):.2f}")
PY
