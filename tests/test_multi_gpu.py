"""N > 1 path on real GPUs: one process per GPU over NCCL; per-shard coupling + all-gather of the
sampled index pairs (cfm_b200/dist.py).  Skipped on boxes with fewer than 2 GPUs; the world_size-2
host logic is covered on CPU with gloo in tests/test_host.py."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    import cfm_b200
    from cfm_b200 import dist as cdist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        n_local, d = 256, 32
        g = torch.Generator().manual_seed(100 + rank)
        x0, x1 = torch.randn(n_local, d, generator=g).to(dev), torch.randn(n_local, d, generator=g).to(dev)
        s = cfm_b200.OTPlanSampler("exact", warn=False)
        np.random.seed(7 + rank)
        a, b, ig, jg = cdist.sharded_sample_plan(s, x0, x1)
        # every rank sees the same global pairing; its own slice indexes its own shard
        lo = rank * n_local
        own_i, own_j = ig[lo:lo + n_local] - lo, jg[lo:lo + n_local] - lo
        ok_own = torch.equal(a, x0[own_i]) and torch.equal(b, x1[own_j])
        # per-shard coupling == single-GPU coupling of the same shard with the same RNG stream
        np.random.seed(7 + rank)
        ra, rb = s.sample_plan(x0, x1)
        ok_same = torch.equal(a, ra) and torch.equal(b, rb)
        sk = cfm_b200.OTPlanSampler("sinkhorn", reg=0.05, normalize_cost=True, num_iter_max=50, stop_thr=0.0, warn=False)
        i2, j2, ig2, jg2 = cdist.sharded_sample_pairs(sk, x0, x1)
        q.put((rank, ok_own, ok_same, ig.cpu().tolist(), jg.cpu().tolist(), int(ig2.numel()),
               bool(((ig2 >= 0) & (ig2 < world * n_local)).all().item())))
    finally:
        dist.destroy_process_group()


def test_sharded_coupling_nccl_two_gpus():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=300) for _ in range(2))
    [p.join(120) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    (r0, own0, same0, ig0, jg0, n0, in0), (r1, own1, same1, ig1, jg1, n1, in1) = res
    assert own0 and own1 and same0 and same1
    assert ig0 == ig1 and jg0 == jg1 and len(ig0) == 512  # identical gathered pairing on both ranks
    # exact OT draws with replacement from a permutation plan: every drawn pair is (i, sigma(i)) of its own shard,
    # so within a shard equal i imply equal j and distinct i imply distinct j (sigma is a bijection)
    for lo in (0, 256):
        pairs = set(zip(ig0[lo:lo + 256], jg0[lo:lo + 256]))
        assert len({i for i, _ in pairs}) == len(pairs) == len({j for _, j in pairs})
        assert all(lo <= i < lo + 256 and lo <= j < lo + 256 for i, j in pairs)  # rank offsets applied to both
    assert n0 == n1 == 512 and in0 and in1


def _ode_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    import cfm_b200
    from cfm_b200 import dist as cdist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        torch.manual_seed(0)
        m = cfm_b200.MLP(dim=64, w=256, time_varying=True).to(dev)
        g = torch.Generator().manual_seed(5)
        x = (torch.randn(2048, 64, generator=g) * torch.linspace(0.2, 3.0, 2048)[:, None]).to(dev)  # shards differ
        span = torch.linspace(0, 1, 3)
        node = cfm_b200.NeuralODE(cfm_b200.torch_wrapper(m), solver="dopri5", atol=1e-5, rtol=1e-5)
        full = node.trajectory(x, span)                     # single-process run on the whole batch
        s_full = dict(node.stats)
        lock = cdist.sharded_trajectory(node, x, span, lockstep=True)
        s_lock = dict(node.stats)
        free = cdist.sharded_trajectory(node, x, span, lockstep=False)
        s_free = dict(node.stats)
        q.put((rank, s_full["nfe"], s_full["accepted"], s_full["rejected"], s_lock["nfe"], s_lock["accepted"],
               s_lock["rejected"], s_free["nfe"], float((lock - full).abs().max()), float((free - full).abs().max()),
               float(full.abs().max()), tuple(lock.shape)))
    finally:
        dist.destroy_process_group()


def test_sharded_ode_lockstep_takes_the_single_process_step_sequence():
    """rows sharded over 2 ranks; with the one-float all-reduce of the error sums both ranks take exactly
    the step sequence of the 1-GPU run on the whole batch (same NFE / accepted / rejected) and reproduce its states to
    rounding; independent controllers agree to the solver tolerance only."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29800 + os.getpid() % 1000
    procs = [ctx.Process(target=_ode_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=600) for _ in range(2))
    [p.join(120) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    for rank, nfe, acc, rej, lnfe, lacc, lrej, fnfe, dlock, dfree, amax, shape in res:
        assert (lnfe, lacc, lrej) == (nfe, acc, rej), res
        assert shape == (3, 2048, 64)
        assert dlock <= 2e-6 * max(1.0, amax), res   # same steps; only the summation order of the norm differs
        assert dfree <= 1e-3 * max(1.0, amax), res
