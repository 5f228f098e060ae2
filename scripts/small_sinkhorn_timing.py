import sys, time, torch, numpy as np
sys.path.insert(0, '.')
import cfm_b200
from oracle import coupling as oc
dev = torch.device('cuda:0')
for n, d in ((128, 2), (256, 2), (512, 16), (1024, 64), (2048, 128)):
    g = torch.Generator().manual_seed(n)
    x0, x1 = torch.randn(n, d, generator=g), torch.randn(n, d, generator=g) + 0.5
    for kw, tag in ((dict(reg=0.05, normalize_cost=True), 'reg=.05 norm'), (dict(reg=0.5), 'reg=.5')):
        s = cfm_b200.OTPlanSampler('sinkhorn', warn=False, **kw)
        a, b = x0.to(dev), x1.to(dev)
        for _ in range(3): s.sample_plan(a, b)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        reps = 20
        for _ in range(reps): s.sample_plan(a, b)
        torch.cuda.synchronize(); gpu = (time.perf_counter() - t0) / reps
        s.warn = True; s.sample_plan(a, b); info = dict(s.last_info); s.warn = False
        s100 = cfm_b200.OTPlanSampler('sinkhorn', warn=False, num_iter_max=100, stop_thr=0.0, **kw)
        for _ in range(3): s100.sample_plan(a, b)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(reps): s100.sample_plan(a, b)
        torch.cuda.synchronize(); gpu100 = (time.perf_counter() - t0) / reps
        t0 = time.perf_counter()
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            for _ in range(2): oc.sample_plan(x0, x1, 'sinkhorn', kw['reg'], kw.get('normalize_cost', False))
        cpu = (time.perf_counter() - t0) / 2
        print(f"n={n} d={d} {tag}: gpu default-stop {gpu*1e3:.3f} ms ({info.get('iterations')} it, precise={info.get('precise')})  gpu 100it {gpu100*1e3:.3f} ms  cpu-oracle {cpu*1e3:.1f} ms")
