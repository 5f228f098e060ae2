"""BASELINE config 4 shard: N=4096, d=784, un-normalised reg=0.1 (|M/reg| ~ 1.5e4 -> float64 potentials), 100 it."""
import sys, time, torch, numpy as np
sys.path.insert(0, '.')
import cfm_b200
dev = torch.device('cuda:0')
torch.manual_seed(0)
N, D = 4096, 784
x0, x1 = torch.randn(N, D, device=dev), torch.randn(N, D, device=dev)
for prec in ("fp64", "fp64-mixed", "auto"):
    s = cfm_b200.OTPlanSampler("sinkhorn", reg=0.1, num_iter_max=100, stop_thr=0.0, warn=False, precision=prec)
    for _ in range(2): s.sample_plan(x0, x1)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): s.sample_plan(x0, x1)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
    s.warn = True; s.sample_plan(x0, x1)
    cp = s._couple(x0, x1, dev)
    P = (-(cp.M[:, :N].double() / 0.1) + cp.log_u[:, None] + cp.log_v[None, :]).exp()
    print(f"precision={prec}: {dt*1e3:.2f} ms/coupling  info={s.last_info}  row marginal rel err {float((P.sum(1)*N-1).abs().max()):.2e} col {float((P.sum(0)*N-1).abs().max()):.2e}")
