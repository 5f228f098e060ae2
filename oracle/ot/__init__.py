"""TEST INFRASTRUCTURE ONLY -- CPU stand-in for the POT package (``import ot``).

The reference (torchcfm) delegates all solver arithmetic of the coupling path to
the un-vendored, unpinned third-party package POT (``import ot as pot``,
/root/reference/torchcfm/optimal_transport.py:7; declared in setup.py:14,
requirements.txt:12).  POT is not installed in this image and cannot be fetched
(no network), so this module restates -- from POT 0.9.x's published algorithms --
exactly the functions the reference calls:

    ot.unif        optimal_transport.py:79, :292
    ot.emd         optimal_transport.py:49  (via functools.partial, called at :87)
    ot.sinkhorn    optimal_transport.py:51  (called at :87)
    ot.emd2        optimal_transport.py:286 (called at :300)
    ot.sinkhorn2   optimal_transport.py:288 (called at :300)

PARITY STATUS
  * emd / emd2: pinned mathematically.  For uniform equal-size marginals the LP
    optimum is P_sigma / N with sigma the optimal assignment; sigma is unique
    almost surely for continuous data and is computed here with
    scipy.optimize.linear_sum_assignment -- an independent exact solver that the
    reference itself uses (optimal_transport.py:179).
  * sinkhorn / sinkhorn2: PARITY UNPINNED at the POT boundary.  No golden vector
    for Sinkhorn exists anywhere in the reference; the bodies below are restated
    from the published POT source (ot/bregman/_sinkhorn.py: sinkhorn_knopp,
    sinkhorn_log) and are anchored only by the analytic 2x2 known-answer test
    (tests/test_oracle.py) and by the reference's own call sites.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import
this package.  Nothing under cfm_b200/ imports it: the product path has no CPU
route.
"""
import warnings

import numpy as np
from scipy.optimize import linear_sum_assignment
from scipy.special import logsumexp as _logsumexp

__version__ = "0.9.x-oracle-shim"


def unif(n, type_as=None):
    """POT ot/utils.py::unif -- uniform histogram, float64."""
    return np.ones((n,)) / n


def _as_f64_c(x):
    return np.asarray(x, dtype=np.float64, order="C")


def emd(a, b, M, numItermax=100000, log=False, center_dual=True, numThreads=1,
        check_marginals=True):
    """POT ot/lp/__init__.py::emd restricted to what the reference feeds it.

    POT casts a, b, M to float64 C-order and runs a network simplex; the result
    is an optimal vertex of the transport polytope.  For a = b = 1/N that vertex
    is a permutation matrix / N.  Uniform marginals of DIFFERENT sizes (what
    OTPlanSampler passes when the two batches differ, optimal_transport.py:79)
    are solved as the transport LP itself with SciPy's HiGHS simplex -- an
    independent exact LP solver; the optimum is unique almost surely for
    continuous data, so it is the vertex POT's network simplex returns.
    Non-uniform marginals are outside the reference hot path and raise.
    """
    a, b, M = _as_f64_c(a), _as_f64_c(b), _as_f64_c(M)
    if a.size == 0:
        a = np.ones((M.shape[0],)) / M.shape[0]
    if b.size == 0:
        b = np.ones((M.shape[1],)) / M.shape[1]
    n0, n1 = M.shape
    if not (np.allclose(a, 1.0 / n0) and np.allclose(b, 1.0 / n1)):
        raise NotImplementedError(
            "oracle ot.emd restates POT only for uniform marginals "
            "(the only case OTPlanSampler produces)")
    if n0 != n1:
        G = _transport_lp(a, b, M)
        if log:
            return G, {"cost": float((G * M).sum()), "warning": None}
        return G
    row, col = linear_sum_assignment(M)
    G = np.zeros((n0, n1), dtype=np.float64)
    G[row, col] = a[row]
    if log:
        return G, {"cost": float((G * M).sum()), "warning": None}
    return G


def _transport_lp(a, b, M):
    """min <G, M>  s.t.  G 1 = a, G^T 1 = b, G >= 0  (dual simplex, vertex solution)."""
    from scipy.optimize import linprog
    from scipy.sparse import kron, eye, csr_matrix
    n0, n1 = M.shape
    A_rows = kron(eye(n0), np.ones((1, n1)))          # row sums
    A_cols = kron(np.ones((1, n0)), eye(n1))          # column sums
    A = csr_matrix(np.vstack([A_rows.toarray(), A_cols.toarray()[:-1]]))  # drop one redundant constraint
    rhs = np.concatenate([a, b[:-1]])
    res = linprog(M.reshape(-1), A_eq=A, b_eq=rhs, bounds=(0, None), method="highs-ds")
    if res.status != 0:
        raise RuntimeError(f"oracle transport LP failed: {res.message}")
    G = res.x.reshape(n0, n1)
    G[np.abs(G) < 1e-15] = 0.0
    return G


def emd2(a, b, M, processes=1, numItermax=100000, log=False, return_matrix=False,
         center_dual=True, numThreads=1, check_marginals=True):
    """POT ot/lp/__init__.py::emd2 -- the LP objective sum(G * M), float."""
    M64 = _as_f64_c(M)
    G = emd(a, b, M64, numItermax=numItermax, numThreads=numThreads)
    return float(np.sum(G * M64))


def sinkhorn_knopp(a, b, M, reg, numItermax=1000, stopThr=1e-9, verbose=False,
                   log=False, warn=True, warmstart=None, **kwargs):
    """POT ot/bregman/_sinkhorn.py::sinkhorn_knopp (kernel-space Sinkhorn-Knopp).

    dtype pattern kept as in POT's NumPy backend: u, v start in M.dtype (fp32
    when M came from a fp32 torch tensor, optimal_transport.py:87), K = exp(M /
    -reg) in M.dtype, a/b float64 so everything is float64 after the first
    half-iteration.  Update order v then u; marginal check every 10 iterations.
    """
    a = np.asarray(a)
    b = np.asarray(b)
    M = np.asarray(M)
    if len(a) == 0:
        a = np.full((M.shape[0],), 1.0 / M.shape[0], dtype=M.dtype)
    if len(b) == 0:
        b = np.full((M.shape[1],), 1.0 / M.shape[1], dtype=M.dtype)
    dim_a, dim_b = len(a), len(b)
    u = np.ones(dim_a, dtype=M.dtype) / dim_a
    v = np.ones(dim_b, dtype=M.dtype) / dim_b
    K = np.exp(M / (-reg))
    Kp = (1 / a).reshape(-1, 1) * K
    err = 1.0
    niter = 0
    for ii in range(numItermax):
        niter = ii
        uprev, vprev = u, v
        KtransposeU = np.dot(K.T, u)
        with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
            v = b / KtransposeU
            u = 1.0 / np.dot(Kp, v)
        if (np.any(KtransposeU == 0) or np.any(np.isnan(u)) or np.any(np.isnan(v))
                or np.any(np.isinf(u)) or np.any(np.isinf(v))):
            warnings.warn("Warning: numerical errors at iteration %d" % ii)
            u, v = uprev, vprev
            break
        if ii % 10 == 0:
            tmp2 = np.einsum("i,ij,j->j", u, K, v)
            err = np.linalg.norm(tmp2 - b)
            if err < stopThr:
                break
    else:
        if warn:
            warnings.warn("Sinkhorn did not converge. You might want to increase "
                          "the number of iterations `numItermax` or the "
                          "regularization parameter `reg`.")
    P = u.reshape((-1, 1)) * K * v.reshape((1, -1))
    if log:
        return P, {"u": u, "v": v, "err": err, "niter": niter}
    return P


def sinkhorn_log(a, b, M, reg, numItermax=1000, stopThr=1e-9, verbose=False,
                 log=False, warn=True, warmstart=None, **kwargs):
    """POT ot/bregman/_sinkhorn.py::sinkhorn_log (log-domain Sinkhorn).

    This is the algorithm BASELINE.json's north_star names ("log-domain Sinkhorn
    iterations").  Mr = -M/reg stays in M.dtype (fp32); loga/logb are float64 so
    the potentials and the exponent sums are float64 from the first update on.
    """
    a = np.asarray(a)
    b = np.asarray(b)
    M = np.asarray(M)
    if len(a) == 0:
        a = np.full((M.shape[0],), 1.0 / M.shape[0], dtype=M.dtype)
    if len(b) == 0:
        b = np.full((M.shape[1],), 1.0 / M.shape[1], dtype=M.dtype)
    dim_a, dim_b = len(a), len(b)
    Mr = -M / reg
    u = np.zeros(dim_a, dtype=M.dtype)
    v = np.zeros(dim_b, dtype=M.dtype)
    loga, logb = np.log(a), np.log(b)
    err = 1.0
    niter = 0
    for ii in range(numItermax):
        niter = ii
        v = logb - _logsumexp(Mr + u[:, None], axis=0)
        u = loga - _logsumexp(Mr + v[None, :], axis=1)
        if ii % 10 == 0:
            tmp2 = np.sum(np.exp(Mr + u[:, None] + v[None, :]), axis=0)
            err = np.linalg.norm(tmp2 - b)
            if err < stopThr:
                break
    else:
        if warn:
            warnings.warn("Sinkhorn did not converge. You might want to increase "
                          "the number of iterations `numItermax` or the "
                          "regularization parameter `reg`.")
    P = np.exp(Mr + u[:, None] + v[None, :])
    if log:
        return P, {"log_u": u, "log_v": v, "err": err, "niter": niter}
    return P


def sinkhorn(a, b, M, reg, method="sinkhorn", numItermax=1000, stopThr=1e-9,
             verbose=False, log=False, warn=True, warmstart=None, **kwargs):
    """POT ot/bregman/_sinkhorn.py::sinkhorn -- dispatch on ``method``.

    The reference never passes ``method`` (optimal_transport.py:51) so it runs
    sinkhorn_knopp; 'sinkhorn_log' is the variant north_star targets.
    """
    m = method.lower()
    if m == "sinkhorn":
        return sinkhorn_knopp(a, b, M, reg, numItermax=numItermax, stopThr=stopThr,
                              verbose=verbose, log=log, warn=warn, warmstart=warmstart)
    if m == "sinkhorn_log":
        return sinkhorn_log(a, b, M, reg, numItermax=numItermax, stopThr=stopThr,
                            verbose=verbose, log=log, warn=warn, warmstart=warmstart)
    raise ValueError("Unknown method '%s'." % method)


def sinkhorn2(a, b, M, reg, method="sinkhorn", numItermax=1000, stopThr=1e-9,
              verbose=False, log=False, warn=False, warmstart=None, **kwargs):
    """POT ot/bregman/_sinkhorn.py::sinkhorn2 -- sum(M * plan) for one histogram."""
    P = sinkhorn(a, b, M, reg, method=method, numItermax=numItermax, stopThr=stopThr,
                 verbose=verbose, log=False, warn=warn, warmstart=warmstart)
    return np.sum(np.asarray(M) * P)


class _Unsupported:
    def __init__(self, name):
        self._name = name

    def __getattr__(self, item):
        def _raise(*a, **k):
            raise NotImplementedError(
                "oracle ot.%s.%s: outside the north_star hot path" % (self._name, item))
        return _raise


# referenced (but never called on the hot path) at optimal_transport.py:53,55
unbalanced = _Unsupported("unbalanced")
partial = _Unsupported("partial")
