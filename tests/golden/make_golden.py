"""Generate the committed golden fixtures from the UNMODIFIED reference package.

Run in the build container only (the GPU box has no /root/reference):

    PYTHONPATH=/root/repo/oracle:/root/reference python tests/golden/make_golden.py

``torchcfm`` is imported verbatim from /root/reference; the only stand-in is the
module named ``ot`` (oracle/ot, the POT shim -- POT itself is not installable
here).  So every array below was produced by the reference's own glue
(cdist**2, flattening, uniform marginals, guards, np.random.choice + divmod,
gather, matcher formulas, torch.nn MLP); the exact-OT solver body is SciPy's LSA
and the Sinkhorn body is the POT restatement (parity unpinned, see oracle/ot).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))

import ot  # noqa: E402  (oracle/ot shim)
assert "oracle" in ot.__version__, "expected the oracle POT shim on PYTHONPATH"
import torchcfm  # noqa: E402
from torchcfm.conditional_flow_matching import (  # noqa: E402
    ConditionalFlowMatcher, ExactOptimalTransportConditionalFlowMatcher,
    SchrodingerBridgeConditionalFlowMatcher, TargetConditionalFlowMatcher,
    VariancePreservingConditionalFlowMatcher)
from torchcfm.models.models import MLP  # noqa: E402
from torchcfm.optimal_transport import OTPlanSampler, wasserstein  # noqa: E402

assert torchcfm.__file__.startswith("/root/reference"), torchcfm.__file__


def eight_gaussians(n, gen):
    """Synthetic stand-in for utils.sample_8gaussians (utils.py:11-41; that module
    needs matplotlib+torchdyn, both absent): 8 centres on a radius-5 circle, per-axis
    std 0.1**0.25."""
    ang = torch.arange(8) * (2 * np.pi / 8)
    centers = 5.0 * torch.stack([torch.cos(ang), torch.sin(ang)], 1)
    k = torch.randint(0, 8, (n,), generator=gen)
    return (centers[k] + (0.1 ** 0.25) * torch.randn(n, 2, generator=gen)).float()


def two_moons(n, gen):
    """Synthetic stand-in for utils.sample_moons (utils.py:35-37): generate_moons(n,
    noise=0.2) * 3 - 1."""
    n_out = n // 2
    n_in = n - n_out
    to = torch.linspace(0, np.pi, n_out)
    ti = torch.linspace(0, np.pi, n_in)
    outer = torch.stack([torch.cos(to), torch.sin(to)], 1)
    inner = torch.stack([1 - torch.cos(ti), 1 - torch.sin(ti) - 0.5], 1)
    x = torch.cat([outer, inner], 0) + 0.2 * torch.randn(n, 2, generator=gen)
    x = x[torch.randperm(n, generator=gen)]
    return (x * 3 - 1).float()


def main():
    out = {}

    # ---- exact OT, the reference test shape (tests/test_optimal_transport.py:44-60)
    torch.manual_seed(1980)
    np.random.seed(1980)
    x0 = torch.randn(128, 2, 2, 2)
    x1 = torch.randn(128, 2, 2, 2)
    s = OTPlanSampler(method="exact")
    pi = s.get_map(x0, x1)
    i, j = s.sample_map(pi, batch_size=128, replace=True)
    out["exact128_x0"], out["exact128_x1"] = x0.numpy(), x1.numpy()
    out["exact128_sigma"] = pi.argmax(1).astype(np.int64)
    assert np.array_equal(pi[np.arange(128), out["exact128_sigma"]], np.full(128, 1 / 128))
    out["exact128_i"], out["exact128_j"] = i, j
    torch.manual_seed(1980)
    np.random.seed(1980)
    _ = torch.randn(128, 2, 2, 2), torch.randn(128, 2, 2, 2)
    sx0, sx1 = s.sample_plan(x0, x1, replace=True)
    assert torch.equal(sx0, x0[i]) and torch.equal(sx1, x1[j])

    # ---- exact OT, BASELINE config 1 shape: 8gaussians -> moons, N=256, d=2
    gen = torch.Generator().manual_seed(7)
    g8, mo = eight_gaussians(256, gen), two_moons(256, gen)
    pi = OTPlanSampler(method="exact").get_map(g8, mo)
    out["c1_x0"], out["c1_x1"] = g8.numpy(), mo.numpy()
    out["c1_sigma"] = pi.argmax(1).astype(np.int64)
    # non power-of-two N and normalize_cost
    gen = torch.Generator().manual_seed(11)
    a100, b100 = torch.randn(100, 5, generator=gen), torch.randn(100, 5, generator=gen) + 1
    pi = OTPlanSampler(method="exact", normalize_cost=True).get_map(a100, b100)
    out["e100_x0"], out["e100_x1"] = a100.numpy(), b100.numpy()
    out["e100_sigma"] = pi.argmax(1).astype(np.int64)
    np.random.seed(5)
    i, j = OTPlanSampler(method="exact").sample_map(pi, 100, replace=True)
    out["e100_i"], out["e100_j"] = i, j

    # ---- Sinkhorn through the reference glue (solver body = POT restatement)
    gen = torch.Generator().manual_seed(3)
    y0, y1 = torch.randn(128, 16, generator=gen), torch.randn(128, 16, generator=gen)
    pi = OTPlanSampler(method="sinkhorn", reg=0.05, normalize_cost=True).get_map(y0, y1)
    out["sk128_x0"], out["sk128_x1"], out["sk128_plan"] = y0.numpy(), y1.numpy(), pi
    np.random.seed(9)
    i, j = OTPlanSampler(method="sinkhorn").sample_map(pi, 128, replace=True)
    out["sk128_i"], out["sk128_j"] = i, j
    # un-normalised, larger reg (kernel-space finite)
    pi = OTPlanSampler(method="sinkhorn", reg=4.0).get_map(y0, y1)
    out["sk128r4_plan"] = pi

    # ---- wasserstein (tests/test_optimal_transport.py:63-91)
    torch.manual_seed(1980)
    w0, w1 = torch.randn(128, 2, 2, 2), torch.randn(128, 2, 2, 2)
    out["w_x0"], out["w_x1"] = w0.numpy(), w1.numpy()
    out["w_vals"] = np.array([wasserstein(w0, w1, "exact"),
                              wasserstein(w0, w1, "exact", power=1),
                              wasserstein(w0, w1, "sinkhorn", reg=0.01, power=1)])

    # ---- the five matchers (tests/test_conditional_flow_matcher.py:93-127)
    torch.manual_seed(21)
    m0, m1 = torch.randn(128, 3, 4), torch.randn(128, 3, 4)
    out["fm_x0"], out["fm_x1"] = m0.numpy(), m1.numpy()
    for name, fm in [("i_cfm", ConditionalFlowMatcher(0.5)),
                     ("exact_ot_cfm", ExactOptimalTransportConditionalFlowMatcher(0.5)),
                     ("t_cfm", TargetConditionalFlowMatcher(0.5)),
                     ("sb_cfm", SchrodingerBridgeConditionalFlowMatcher(0.5, ot_method="exact")),
                     ("vp_cfm", VariancePreservingConditionalFlowMatcher(0.5))]:
        torch.manual_seed(1994)
        np.random.seed(1994)
        t, xt, ut, eps = fm.sample_location_and_conditional_flow(m0, m1, return_noise=True)
        out[f"fm_{name}_t"], out[f"fm_{name}_xt"] = t.numpy(), xt.numpy()
        out[f"fm_{name}_ut"], out[f"fm_{name}_eps"] = ut.numpy(), eps.numpy()
        out[f"fm_{name}_lambda"] = np.asarray(fm.compute_lambda(t))

    # ---- reference MLP forward (models.py:4-21) -- pure torch, fully pinned
    torch.manual_seed(0)
    mlp = MLP(dim=16, w=64, time_varying=True)
    xin = torch.randn(64, 17)
    for k, v in mlp.state_dict().items():
        out[f"mlp16_{k}"] = v.numpy()
    out["mlp16_x"] = xin.numpy()
    with torch.no_grad():
        out["mlp16_y"] = mlp(xin).numpy()
        out["mlp16_y64"] = mlp.double()(xin.double()).numpy()
    # BASELINE config 3 architecture: weights regenerated from the seed by the test
    torch.manual_seed(0)
    big = MLP(dim=784, w=256, time_varying=True)
    xb = torch.randn(8, 785)
    out["mlp784_x"] = xb.numpy()
    with torch.no_grad():
        out["mlp784_y"] = big(xb).numpy()
        out["mlp784_y64"] = big.double()(xb.double()).numpy()
    out["mlp784_nparams"] = np.array(sum(p.numel() for p in big.parameters()))

    path = os.path.join(HERE, "reference_vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", len(out), "arrays")


if __name__ == "__main__":
    sys.exit(main())
