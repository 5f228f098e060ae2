"""BASELINE config 4 (N = 4096 per shard, d = 512, eps = 0.1): how far the Sinkhorn potentials move per iteration and how
many terms survive the seeded screening of the float64-potential solver (threshold = previous log-sum-exp + global
minimum change of the other potential - 34).  DESIGN.md section 3.2.  Needs ~1 GB of RAM and a few minutes.

    python scripts/sim/sinkhorn_screening.py
"""
import numpy as np
import torch
from scipy.special import logsumexp

g = torch.Generator().manual_seed(40)
N, d = 4096, 512
x0, x1 = torch.randn(N, d, generator=g), torch.randn(N, d, generator=g)
M = (torch.cdist(x0, x1) ** 2).float().numpy()
Mr = (-M / np.float32(0.1)).astype(np.float32).astype(np.float64)
u = np.zeros(N); v = -np.log(N) - logsumexp(Mr, axis=0)
for it in range(1, 101):
    un = -np.log(N) - logsumexp(Mr + v[None, :], axis=1)
    du = un - u; u = un
    Y = Mr + u[:, None]
    lse_c = logsumexp(Y, axis=0)
    vn = -np.log(N) - lse_c
    dv = vn - v
    if it in (1,2,3,5,8,12,20,30,50,75,100):
        # hits if seeded: columns: entries with Y_ij > lse_prev_j + du_min - 34, where lse_prev_j = -logN - v_old_j
        thr = (-np.log(N) - v) + du.min() - 34
        hits_c = (Y > thr[None, :]).sum()
        print(f'it {it}: du min {du.min():9.2f} max {du.max():8.2f} | dv min {dv.min():9.2f} max {dv.max():8.2f} | col-phase seeded hits {hits_c} ({hits_c/N/N*100:.3f}% of elements)')
    v = vn
