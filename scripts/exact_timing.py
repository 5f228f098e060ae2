import sys, time, torch, numpy as np
sys.path.insert(0, '.')
import cfm_b200
from oracle import coupling as oc
dev = torch.device('cuda:0')
sizes = ((128, 8), (256, 2), (512, 16), (1024, 32)) if '--small' in sys.argv else ((128, 8), (256, 2), (512, 16), (1024, 32), (2048, 64), (4096, 64))
for n, d in sizes:
    g = torch.Generator().manual_seed(n)
    x0, x1 = torch.randn(n, d, generator=g), torch.randn(n, d, generator=g)
    s = cfm_b200.OTPlanSampler('exact', warn=False)
    a, b = x0.to(dev), x1.to(dev)
    for _ in range(2): s.sample_plan(a, b)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    reps = 20 if n <= 1024 else 3
    for _ in range(reps): s.sample_plan(a, b)
    torch.cuda.synchronize(); gpu = (time.perf_counter() - t0) / reps
    t0 = time.perf_counter()
    for _ in range(3): oc.sample_plan(x0, x1, 'exact')
    cpu = (time.perf_counter() - t0) / 3
    s.warn = True
    sig = s.get_map(a, b).argmax(1)
    cp = s._couple(a, b, dev)
    st = cp.status.cpu().tolist()
    ok = np.array_equal(sig, oc.assignment(oc.cost_matrix(x0, x1)))
    print(f"n={n} d={d} gpu {gpu*1e3:.3f} ms/coupling  cpu-oracle {cpu*1e3:.3f} ms  sigma_exact={ok} augmentations={st[1]} dijkstra_steps={st[2]}")
