"""Golden vectors for OTPlanSampler.sample_trajectory / sample_plan_with_scipy, produced by the
UNMODIFIED reference (torchcfm/optimal_transport.py:147-182, :221-251) on the oracle POT shim.

    PYTHONPATH=/root/repo/oracle:/root/reference python tests/golden/make_golden_trajectory.py

Exact OT: every plan row is one-hot, so the chain is sigma_t composed -- bit-exact gate.
Sinkhorn: the per-row categorical draws depend on the plan values (POT restatement, parity
unpinned): the GPU test gates on the fraction of identical chain entries.
"""
import os

import numpy as np
import torch

import ot  # noqa: F401  (oracle/ot shim)
assert "oracle" in ot.__version__
import torchcfm  # noqa: E402
from torchcfm.optimal_transport import OTPlanSampler  # noqa: E402

assert torchcfm.__file__.startswith("/root/reference"), torchcfm.__file__
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    out = {}
    g = torch.Generator().manual_seed(321)
    n, times, d = 96, 4, 3
    X = torch.randn(n, times, d, generator=g) + torch.arange(times)[None, :, None] * 0.7
    out["traj_X"] = X.numpy()
    np.random.seed(17)
    out["traj_exact"] = np.asarray(OTPlanSampler("exact").sample_trajectory(X))
    np.random.seed(17)
    out["traj_sinkhorn"] = np.asarray(OTPlanSampler("sinkhorn", reg=0.5).sample_trajectory(X))
    np.random.seed(17)
    out["traj_sinkhorn_norm"] = np.asarray(
        OTPlanSampler("sinkhorn", reg=0.1, normalize_cost=True).sample_trajectory(X))
    a, b = X[:, 0], X[:, 1]
    s0, s1 = OTPlanSampler("exact").sample_plan_with_scipy(a, b)
    out["scipy_x0"], out["scipy_x1"] = s0.numpy(), s1.numpy()
    path = os.path.join(HERE, "trajectory_vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
