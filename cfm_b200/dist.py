"""Multi-GPU coupling: minibatch sharding, one process per GPU (torch.distributed).

The reference's only parallelism is data parallelism in which every rank runs its OWN minibatch
OT on its local ``batch_size // world_size`` shard (examples/images/cifar10/train_cifar10_ddp.py
:71-77,169): the coupling never crosses ranks.  BASELINE.json's north_star prescribes the same
partitioning ("coupling stays per-shard, NCCL over NVLink only to gather sampled indices").

``sharded_sample_pairs`` therefore solves the local (N/G x N/G) coupling with the device kernels
and all-gathers the 2 x N/G int64 global indices (<= 128 KB per rank at N = 64k): a latency-bound
NCCL all_gather on a side stream, overlappable with the next coupling.  There is no data-path
collective besides it.
"""
import torch
import torch.distributed as dist


def shard_bounds(n, world_size, rank):
    """Contiguous shard [lo, hi) of ``n`` rows for ``rank`` (first n % world_size ranks get +1)."""
    base, rem = divmod(n, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def sharded_sample_pairs(sampler, x0_local, x1_local, group=None, gather=True, equal_shards=True,
                         pair_fn=None):
    """Per-shard coupling + optional all-gather of the sampled index pairs.

    x0_local / x1_local: this rank's shard.  Returns (i_local, j_local) device index tensors into
    the shard and, if ``gather``, also (i_global, j_global): the concatenation over ranks of the
    pairs offset into the global batch.  ``equal_shards`` (the DDP case: batch_size // world_size
    rows everywhere, train_cifar10_ddp.py:74) makes rank r's rows start at r * n_local and needs ONE
    collective per coupling; ragged shards first all-gather the shard sizes.
    ``pair_fn`` (tests only) replaces ``sampler.sample_pairs``.
    """
    fn = pair_fn if pair_fn is not None else sampler.sample_pairs
    i, j = fn(x0_local, x1_local)
    if not gather or not dist.is_available() or not dist.is_initialized():
        return (i, j, i, j) if gather else (i, j)
    ws = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n_local = x0_local.shape[0]
    if equal_shards:
        sizes = [n_local] * ws
    else:
        nl = torch.tensor([n_local], dtype=torch.int64, device=i.device)
        got = [torch.zeros_like(nl) for _ in range(ws)]
        dist.all_gather(got, nl, group=group)
        sizes = [int(s.item()) for s in got]
    off = sum(sizes[:rank])
    pairs = torch.stack([i + off, j + off])  # (2, n_local) int64, global row numbers
    pad = max(sizes)
    if pairs.shape[1] != pad:  # ragged shards: pad to the largest so every rank sends equal bytes
        buf = torch.zeros((2, pad), dtype=pairs.dtype, device=pairs.device)
        buf[:, :pairs.shape[1]] = pairs
        pairs = buf
    outs = [torch.empty_like(pairs) for _ in range(ws)]
    dist.all_gather(outs, pairs.contiguous(), group=group)
    chunks = [o[:, :n] for o, n in zip(outs, sizes)]
    ig = torch.cat([c[0] for c in chunks])
    jg = torch.cat([c[1] for c in chunks])
    return i, j, ig, jg


def sharded_sample_plan(sampler, x0_local, x1_local, group=None):
    """Per-shard ``sample_plan`` (what every DDP rank of the reference does) plus the all-gather of
    the global index pairs.  Returns (x0_local[i], x1_local[j], i_global, j_global)."""
    i, j, ig, jg = sharded_sample_pairs(sampler, x0_local, x1_local, group=group, gather=True)
    return sampler._gather(x0_local, i), sampler._gather(x1_local, j), ig, jg


def sharded_trajectory(node, x, t_span, group=None, gather=True, integrate_fn=None):
    """ODE sampling across ranks ("MLP / ODE sampling"): rows are independent and the weights
    are replicated, so rank r integrates its contiguous shard of ``x`` with its own step controller -- what a
    data-parallel caller of torchdyn does, zero communication during the integration -- and, if ``gather``, one
    all_gather returns the full (len(t_span), B, *dim) trajectory on every rank.

    Per-shard controllers take the step sequence their own shard's error norm dictates (torchdyn's norm is a
    mean over whatever batch it is given), so the result agrees with a single-process run to the solver
    tolerance, not bit for bit.  ``integrate_fn`` (tests only) replaces ``node.trajectory``.
    """
    fn = integrate_fn if integrate_fn is not None else node.trajectory
    if not dist.is_available() or not dist.is_initialized():
        return fn(x, t_span)
    ws, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = shard_bounds(x.shape[0], ws, rank)
    local = fn(x[lo:hi], t_span)  # (T, hi - lo, *dim)
    if not gather:
        return local
    sizes = [shard_bounds(x.shape[0], ws, r) for r in range(ws)]
    pad = max(h - l for l, h in sizes)
    buf = local.new_zeros((local.shape[0], pad) + tuple(local.shape[2:]))
    buf[:, :hi - lo] = local
    outs = [torch.empty_like(buf) for _ in range(ws)]
    dist.all_gather(outs, buf.contiguous(), group=group)
    return torch.cat([o[:, :h - l] for o, (l, h) in zip(outs, sizes)], dim=1)
