"""Per-CTA timeline of the tcgen05 GEMM kernel (debug aid): launch one MLP forward / one cost matrix
with cfm_tc_debug_buffer set and print the checkpoint deltas."""
import sys
import torch
sys.path.insert(0, '.')
import cfm_b200
from cfm_b200 import _ffi
from cfm_b200.optimal_transport import OTPlanSampler
L = _ffi.lib()
dev = torch.device('cuda:0')
dbg = torch.zeros(8 * 148, dtype=torch.int64, device=dev)

def report(tag):
    torch.cuda.synchronize()
    d = dbg.cpu().view(148, 8)
    act = d[:, 0] > 0
    if act.sum() == 0:
        print(tag, 'no data'); return
    t0 = d[act, 0].min()
    rel = (d[act] - t0).float() / 1e3
    names = ['entry', 'alloc', 'first_full', 'acc_ready', 'epi1_done', 'epi_all']
    print(tag, 'CTAs', int(act.sum()), ' '.join(f"{n}:{rel[:, k].median():.1f}/{rel[:, k].max():.1f}us" for k, n in enumerate(names)))
    dbg.zero_()

torch.manual_seed(0)
m = cfm_b200.MLP(dim=784, w=256, time_varying=True).to(dev)
m.mlp_algo = 2
x = torch.randn(10000, 784, device=dev)
with torch.no_grad():
    for _ in range(3): m.vector_field(0.3, x)
    torch.cuda.synchronize()
    # one layer at a time is not exposed; timeline of the LAST layer launch of a forward (4 launches overwrite)
    _ffi.check(L.cfm_tc_debug_buffer(_ffi.ptr(dbg)), 'dbg')
    m.vector_field(0.3, x)
    report('mlp forward (last layer written last)')
s = OTPlanSampler('sinkhorn', cost_algo=2)
x0, x1 = torch.randn(8192, 784, device=dev), torch.randn(8192, 784, device=dev)
for _ in range(2): s._cost(x0, x1, dev)
torch.cuda.synchronize(); dbg.zero_()
s._cost(x0, x1, dev)
report('sqdist 8192x8192x784')
# small-K single layer shape through the cost kernel: 10000 x 256, K=256
x0, x1 = torch.randn(10000, 256, device=dev), torch.randn(256, 256, device=dev)
for _ in range(2): s._cost(x0, x1, dev)
torch.cuda.synchronize(); dbg.zero_()
s._cost(x0, x1, dev)
report('sqdist 10000x256x256 (one tile per CTA)')
_ffi.check(L.cfm_tc_debug_buffer(None), 'dbg')
