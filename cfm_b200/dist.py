"""Multi-GPU coupling and sampling: minibatch sharding, one process per GPU (torch.distributed).

The reference's only parallelism is data parallelism in which every rank runs its OWN minibatch
OT on its local ``batch_size // world_size`` shard (examples/images/cifar10/train_cifar10_ddp.py
:71-77,169): the coupling never crosses ranks.  BASELINE.json's north_star prescribes the same
partitioning ("coupling stays per-shard, NCCL over NVLink only to gather sampled indices").

``sharded_sample_pairs`` therefore solves the local (N/G x N/G) coupling with the device kernels
and all-gathers the 2 x N/G int64 global indices (<= 128 KB per rank at N = 64k).  The collective is
latency-bound, so it is ONE ``all_gather_into_tensor`` into a preallocated (G, 2, n) buffer issued on a
side stream behind an event: the compute stream never waits for it and the next coupling overlaps it;
the consumer synchronises through ``PairGather.wait()`` (or any later use on the compute stream after
``wait``).  There is no data-path collective besides it.

``sharded_trajectory`` shards the ROWS of an ODE batch (weights replicated).  Default: independent
per-shard step controllers (zero communication during the integration, what a data-parallel torchdyn
caller gets).  ``lockstep=True``: the per-shard error sums are all-reduced (one float64 per step attempt)
so that every rank takes exactly the step sequence of a single-process run on the whole batch.
"""
import torch
import torch.distributed as dist


def shard_bounds(n, world_size, rank):
    """Contiguous shard [lo, hi) of ``n`` rows for ``rank`` (first n % world_size ranks get +1)."""
    base, rem = divmod(n, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class PairGather:
    """Handle of an in-flight index all-gather: ``wait()`` makes the current stream wait for it (no host
    block) and returns (i_global, j_global)."""

    def __init__(self, buf, sizes0, event, stream):
        self._buf, self._sizes, self._event, self._stream = buf, sizes0, event, stream
        self._out = None

    def wait(self):
        if self._out is None:
            if self._event is not None:
                torch.cuda.current_stream(self._buf.device).wait_event(self._event)
            buf, sizes = self._buf, self._sizes
            if all(s == buf.shape[2] for s in sizes):  # equal shards: the buffer already is the concatenation
                self._out = (buf[:, 0, :].reshape(-1), buf[:, 1, :].reshape(-1))
            else:
                self._out = (torch.cat([buf[r, 0, :n] for r, n in enumerate(sizes)]),
                             torch.cat([buf[r, 1, :n] for r, n in enumerate(sizes)]))
        return self._out


_side_streams = {}
_gather_bufs = {}


def _side_stream(device):
    s = _side_streams.get(device)
    if s is None:
        s = _side_streams[device] = torch.cuda.Stream(device)
    return s


def sharded_sample_pairs(sampler, x0_local, x1_local, group=None, gather=True, equal_shards=True,
                         pair_fn=None, async_gather=False):
    """Per-shard coupling + optional all-gather of the sampled index pairs.

    x0_local / x1_local: this rank's shard.  Returns (i_local, j_local) device index tensors into
    the shard and, if ``gather``, also (i_global, j_global): the concatenation over ranks of the
    pairs offset into the global batch (i by the x0 shard sizes, j by the x1 shard sizes).
    ``equal_shards`` (the DDP case: batch_size // world_size rows everywhere, train_cifar10_ddp.py:74)
    makes rank r's rows start at r * n_local and needs ONE collective per coupling; ragged shards first
    all-gather the shard sizes.  ``async_gather``: return (i, j, PairGather) instead -- the collective
    runs on a side stream and the caller picks the result up with ``.wait()`` when it needs it.
    ``pair_fn`` (tests only) replaces ``sampler.sample_pairs``.
    """
    fn = pair_fn if pair_fn is not None else sampler.sample_pairs
    i, j = fn(x0_local, x1_local)
    if not gather:
        return i, j
    if not dist.is_available() or not dist.is_initialized():
        if async_gather:
            buf = torch.stack([i, j]).unsqueeze(0)
            return i, j, PairGather(buf, [i.shape[0]], None, None)
        return i, j, i, j
    ws = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n0, n1, nd = x0_local.shape[0], x1_local.shape[0], i.shape[0]
    if equal_shards:
        sizes0, sizes1, sizesd = [n0] * ws, [n1] * ws, [nd] * ws
    else:
        nl = torch.tensor([n0, n1, nd], dtype=torch.int64, device=i.device)
        got = torch.empty((ws, 3), dtype=torch.int64, device=i.device)
        dist.all_gather_into_tensor(got.view(-1), nl, group=group)
        got = got.cpu().tolist()
        sizes0, sizes1, sizesd = [g[0] for g in got], [g[1] for g in got], [g[2] for g in got]
    off0, off1 = sum(sizes0[:rank]), sum(sizes1[:rank])
    pad = max(sizesd)
    dev = i.device
    # preallocated (G, 2, pad) receive buffers per (device, world size, pad), two in rotation so that the
    # gather of coupling k can still be in flight while coupling k+1 fills the other one
    key = (str(dev), ws, pad)
    ring = _gather_bufs.get(key)
    if ring is None:
        ring = _gather_bufs[key] = {"k": 0, "send": [torch.zeros((2, pad), dtype=torch.int64, device=dev) for _ in range(2)],
                                    "recv": [torch.empty((ws, 2, pad), dtype=torch.int64, device=dev) for _ in range(2)]}
    k = ring["k"]
    ring["k"] = k ^ 1
    send, recv = ring["send"][k], ring["recv"][k]
    if dev.type != "cuda":  # gloo (CPU tests): same collective, no streams
        send[0, :nd] = i + off0
        send[1, :nd] = j + off1
        dist.all_gather_into_tensor(recv.view(ws * 2, pad), send, group=group)  # concatenation along dim 0
        h = PairGather(recv, sizesd, None, None)
        return (i, j, h) if async_gather else (i, j) + h.wait()
    cur = torch.cuda.current_stream(dev)
    side = _side_stream(dev)
    side.wait_stream(cur)  # the index tensors are produced on the compute stream
    with torch.cuda.stream(side):
        torch.add(i, off0, out=send[0, :nd])
        torch.add(j, off1, out=send[1, :nd])
        dist.all_gather_into_tensor(recv.view(ws * 2, pad), send, group=group)  # concatenation along dim 0
        ev = torch.cuda.Event()
        ev.record(side)
    i.record_stream(side)
    j.record_stream(side)
    h = PairGather(recv, sizesd, ev, side)
    if async_gather:
        return i, j, h
    ig, jg = h.wait()
    return i, j, ig, jg


def sharded_sample_plan(sampler, x0_local, x1_local, group=None, async_gather=False):
    """Per-shard ``sample_plan`` (what every DDP rank of the reference does) plus the all-gather of
    the global index pairs.  Returns (x0_local[i], x1_local[j], i_global, j_global), or with
    ``async_gather`` (x0_local[i], x1_local[j], PairGather): the collective overlaps the gathers and
    whatever the caller enqueues next."""
    out = sharded_sample_pairs(sampler, x0_local, x1_local, group=group, gather=True, async_gather=async_gather)
    i, j = out[0], out[1]
    a, b = sampler._gather(x0_local, i), sampler._gather(x1_local, j)
    if async_gather:
        return a, b, out[2]
    return a, b, out[2], out[3]


def sharded_trajectory(node, x, t_span, group=None, gather=True, integrate_fn=None, lockstep=False):
    """ODE sampling across ranks ("MLP / ODE sampling"): rows are independent and the weights
    are replicated, so rank r integrates its contiguous shard of ``x`` and, if ``gather``, one all_gather returns
    the full (len(t_span), B, *dim) trajectory on every rank.

    Default: every rank runs its own step controller -- what a data-parallel caller of torchdyn does, zero
    communication during the integration; the result agrees with a single-process run to the solver tolerance
    (torchdyn's error norm is a mean over whatever batch it is given).  ``lockstep=True``: the per-shard sums of
    squares behind that norm are all-reduced (one float64 per step attempt, three more for the initial step), so
    all ranks take exactly the step sequence of a single-process run on the whole batch (same NFE, same accepted /
    rejected steps).  ``integrate_fn`` (tests only) replaces ``node.trajectory``.
    """
    fn = integrate_fn if integrate_fn is not None else node.trajectory
    if not dist.is_available() or not dist.is_initialized():
        return fn(x, t_span)
    ws, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = shard_bounds(x.shape[0], ws, rank)
    if lockstep and integrate_fn is None:
        prev = node.lockstep
        node.lockstep = group if group is not None else True
        try:
            local = fn(x[lo:hi], t_span)  # (T, hi - lo, *dim)
        finally:
            node.lockstep = prev
    else:
        local = fn(x[lo:hi], t_span)
    if not gather:
        return local
    sizes = [shard_bounds(x.shape[0], ws, r) for r in range(ws)]
    pad = max(h - l for l, h in sizes)
    T = local.shape[0]
    # (G, T, pad, *dim) receive buffer, one collective; ragged shards are padded to the largest
    send = local if hi - lo == pad else torch.cat(
        [local, local.new_zeros((T, pad - (hi - lo)) + tuple(local.shape[2:]))], dim=1)
    recv = local.new_empty((ws, T, pad) + tuple(local.shape[2:]))
    dist.all_gather_into_tensor(recv.view((ws * T, pad) + tuple(local.shape[2:])), send.contiguous(), group=group)
    return torch.cat([recv[r, :, :h - l] for r, (l, h) in enumerate(sizes)], dim=1)
