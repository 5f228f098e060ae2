import sys, time, torch, numpy as np
sys.path.insert(0, '.')
import cfm_b200
from cfm_b200 import CouplingStream
dev = torch.device('cuda:0')
N, D = 8192, 784
torch.manual_seed(0); np.random.seed(0)
x0_h = torch.randn(N, D).pin_memory(); x1_h = torch.randn(N, D).pin_memory()
s = cfm_b200.OTPlanSampler("sinkhorn", reg=0.05, normalize_cost=True, num_iter_max=100, stop_thr=0.0, warn=False)
for _ in range(2): s.sample_plan(x0_h, x1_h)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10): s.sample_plan(x0_h, x1_h)
torch.cuda.synchronize()
print('blocking ms/step', (time.perf_counter() - t0) * 100)
for depth in (2, 3):
    pipe = CouplingStream(s, dev, depth=depth)
    for _ in pipe.map((x0_h, x1_h) for _ in range(4)): pass
    torch.cuda.synchronize()
    marks = []
    t0 = time.perf_counter()
    k = 0
    def gen():
        for _ in range(20):
            marks.append(('submit', time.perf_counter() - t0)); yield (x0_h, x1_h)
    for a, b in pipe.map(gen()):
        marks.append(('got', time.perf_counter() - t0))
    torch.cuda.synchronize()
    tot = time.perf_counter() - t0
    print('depth', depth, 'ms/step', tot * 1e3 / 20)
    got = [t for k_, t in marks if k_ == 'got']
    print('  inter-result ms', ' '.join(f"{(b - a) * 1e3:.1f}" for a, b in zip(got[:-1], got[1:])))
# raw copy bandwidth
a_d = torch.empty(N, D, device=dev); h = torch.empty(N, D).pin_memory()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10): a_d.copy_(x0_h, non_blocking=True)
torch.cuda.synchronize(); print('H2D GB/s', 10 * N * D * 4 / (time.perf_counter() - t0) / 1e9)
t0 = time.perf_counter()
for _ in range(10): h.copy_(a_d, non_blocking=True)
torch.cuda.synchronize(); print('D2H GB/s', 10 * N * D * 4 / (time.perf_counter() - t0) / 1e9)
t0 = time.perf_counter()
for _ in range(10): hh = torch.empty(N, D, pin_memory=True)
print('pinned alloc ms', (time.perf_counter() - t0) * 100)
