import sys, torch
sys.path.insert(0, '.')
import cfm_b200
dev = torch.device('cuda:0')
out = []
for n, d in ((128, 128), (256, 2), (512, 128), (1024, 128), (2048, 128)):
    g = torch.Generator().manual_seed(n)
    a, b = torch.randn(n, d, generator=g).to(dev), torch.randn(n, d, generator=g).to(dev)
    s = cfm_b200.OTPlanSampler('sinkhorn', reg=0.05, normalize_cost=True, warn=False, num_iter_max=100, stop_thr=0.0)
    for _ in range(5): s.sample_plan(a, b)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): s.sample_plan(a, b)
    e1.record(); torch.cuda.synchronize()
    out.append(f"n={n}: {e0.elapsed_time(e1) / 20:.3f} ms")
import os
print("CFM_SK_MINROWS =", os.environ.get("CFM_SK_MINROWS", "(default)"), " | ", ", ".join(out))
