"""Per-CTA globaltimer timeline of the fused four-layer MLP kernel and of the fp16x3 cost GEMM (debug aid;
cfm_tc_debug_buffer).  Prints, relative to the kernel's earliest entry, the median / max over CTAs of every
checkpoint: for the fused kernel entry, setup, first operands, last MMA issued of layers 1-4 and, per tile,
"accumulators ready" and "tile drained"."""
import sys
import torch
sys.path.insert(0, '.')
import cfm_b200
from cfm_b200 import _ffi
from cfm_b200.optimal_transport import OTPlanSampler
L = _ffi.lib()
dev = torch.device('cuda:0')
dbg = torch.zeros(64 * 148, dtype=torch.int64, device=dev)


def report(tag, names):
    torch.cuda.synchronize()
    d = dbg.cpu().view(148, 64)
    act = d[:, 0] > 0
    if act.sum() == 0:
        print(tag, 'no data'); return
    t0 = d[act, 0].min()
    print(tag, 'CTAs', int(act.sum()))
    for k, n in names:
        col = d[act, k]
        ok = col > 0
        if ok.sum() == 0:
            continue
        rel = (col[ok] - t0).float() / 1e3
        print(f"   {n:<34} median {rel.median():8.2f} us   max {rel.max():8.2f} us")
    dbg.zero_()


B = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
torch.manual_seed(0)
m = cfm_b200.MLP(dim=784, w=256, time_varying=True).to(dev)
x = torch.randn(B, 784, device=dev)
names = [(0, 'entry'), (1, 'setup done'), (2, 'L1 first operands'), (3, 'L1 last MMA issued'), (4, 'L2 last MMA issued'),
         (5, 'L3 last MMA issued'), (6, 'L4 last MMA issued')]
for tl in range(13):
    lay = tl // 2 + 1 if tl < 6 else 4
    n = tl % 2 if tl < 6 else tl - 6
    names += [(8 + 2 * tl, f'L{lay} tile {n} acc ready'), (9 + 2 * tl, f'L{lay} tile {n} drained')]
names += [(40, 'exit')]
for base, tag in ((42, 'L1 t0'), (48, 'L2 t0'), (54, 'L4 t0')):
    names += [(base, tag + ' warp 2 TMEM loads landed'), (base + 1, tag + ' warp 17 past tfull'),
              (base + 2, tag + ' warp 17 TMEM loads landed'), (base + 3, tag + ' warp 17 done'), (base + 4, tag + ' warp 9 done')]
with torch.no_grad():
    for _ in range(3):
        m.vector_field(0.3, x)
    torch.cuda.synchronize()
    _ffi.check(L.cfm_tc_debug_buffer(_ffi.ptr(dbg)), 'dbg')
    dbg.zero_()
    m.vector_field(0.3, x)
    report(f'fused MLP forward B={B}', names)
    # events around 20 forwards (device time per forward, host launch overhead hidden by the queue)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    _ffi.check(L.cfm_tc_debug_buffer(None), 'dbg')
    y = torch.empty(B, 784, device=dev)
    for _ in range(3):
        m.vector_field(0.3, x, out=y)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20):
        m.vector_field(0.3, x, out=y)
    e1.record()
    torch.cuda.synchronize()
    print('forward incl. input split, 20 back-to-back calls: %.1f us each' % (e0.elapsed_time(e1) * 1e3 / 20))
    import time
    t0 = time.perf_counter()
    for _ in range(20):
        m.vector_field(0.3, x, out=y)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print('host time to enqueue one forward: %.1f us' % ((t1 - t0) * 1e6 / 20))

_ffi.check(L.cfm_tc_debug_buffer(_ffi.ptr(dbg)), 'dbg')
s = OTPlanSampler('sinkhorn')
x0, x1 = torch.randn(8192, 784, device=dev), torch.randn(8192, 784, device=dev)
for _ in range(2):
    s._cost(x0, x1, dev)
torch.cuda.synchronize(); dbg.zero_()
s._cost(x0, x1, dev)
gn = [(0, 'entry'), (1, 'alloc'), (2, 'first_full'), (3, 'acc_ready'), (4, 'epi1_done'), (5, 'epi_all')]
dd = dbg.cpu().view(148, 64)  # the GEMM kernels use 8 slots per CTA: re-index
torch.cuda.synchronize()
d8 = dbg.cpu()[:8 * 148].view(148, 8)
act = d8[:, 0] > 0
t0 = d8[act, 0].min()
rel = (d8[act] - t0).float() / 1e3
print('sqdist fp16x3 8192x8192x784:', ' '.join(f"{n}:{rel[:, k].median():.1f}/{rel[:, k].max():.1f}us" for k, n in gn))
_ffi.check(L.cfm_tc_debug_buffer(None), 'dbg')
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    s._cost(x0, x1, dev)
e1.record(); torch.cuda.synchronize()
print('cost stage (2 pre-passes + GEMM): %.1f us' % (e0.elapsed_time(e1) * 1e3 / 10))
