"""RK-mode fused MLP (cfm_mlp_forward_rkstage_f32): per-stage device time against the two-launch form, and the per-CTA
timeline of layer 1 (first operands / last layer-1 MMA issued / exit).  CFM_RK_PF selects the L2 prefetch distance."""
import os, sys
import torch
sys.path.insert(0, '.')
import cfm_b200
from cfm_b200 import _ffi
L = _ffi.lib()
dev = torch.device('cuda:0')
B, D = 10000, 784
torch.manual_seed(0)
m = cfm_b200.MLP(dim=D, w=256, time_varying=True).to(dev)
x = torch.randn(B, D, device=dev)
k = torch.randn(7, B, D, device=dev) * 0.1
xnew, errp = torch.empty(B, D, device=dev), torch.empty(B, D, device=dev)
hi, lo = torch.empty(B, D, dtype=torch.float16, device=dev), torch.empty(B, D, dtype=torch.float16, device=dev)
tst = torch.zeros(1, device=dev)
st = _ffi.RkState()
st.t, st.dt, st.t_end, st.atol, st.rtol = 0.25, 0.125, 1.0, 1e-4, 1e-4
std = torch.frombuffer(bytearray(bytes(st)), dtype=torch.uint8).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
sp = _ffi.stream_ptr(dev)
dbg = torch.zeros(64 * 148, dtype=torch.int64, device=dev)


def fused(s):
    m.vector_field_rkstage(std, x, k, s, xnew if s == 6 else None, errp if s == 6 else None)


def pair(s):
    _ffi.check(L.cfm_rk_stage_input(_ffi.ptr(std), _ffi.ptr(x), _ffi.ptr(k), _ffi.ptr(xnew if s == 6 else None),
                                    _ffi.ptr(hi), _ffi.ptr(lo), _ffi.ptr(tst), _ffi.ptr(errp if s == 6 else None),
                                    B * D, s, sp), "stage_input")
    m.vector_field_split(tst, hi, lo, k[s])


def timed(fn, s, reps=10):
    for _ in range(2):
        fn(s)
    tot = 0.0
    for _ in range(reps):
        flush.fill_(1)  # cold L2, as inside a step whose working set is 2.5x the L2
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(s); b.record(); torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / reps * 1e3


print('CFM_RK_PF =', os.environ.get('CFM_RK_PF', '(default)'))
with torch.no_grad():
    tf = tp = 0.0
    for s in range(1, 7):
        f, p = timed(fused, s), timed(pair, s)
        tf += f; tp += p
        print(f'stage {s}: one launch {f:7.1f} us   stage-input + MLP {p:7.1f} us')
    print(f'sum over the six stages: one launch {tf:7.1f} us   two launches {tp:7.1f} us')
    _ffi.check(L.cfm_tc_debug_buffer(_ffi.ptr(dbg)), 'dbg')
    for s in (1, 6):
        dbg.zero_(); flush.fill_(1); fused(s); torch.cuda.synchronize()
        d = dbg.cpu().view(148, 64)
        act = d[:, 0] > 0
        t0 = d[act, 0].min()
        out = []
        for slot, name in ((1, 'setup'), (2, 'L1 first operands'), (3, 'L1 last MMA issued'), (4, 'L2 done'), (6, 'L4 issued'), (40, 'exit')):
            rel = (d[act, slot] - t0).float() / 1e3
            out.append(f'{name} {rel.median():.1f}/{rel.max():.1f}')
        print(f'stage {s} timeline (median/max us over {int(act.sum())} CTAs): ' + ', '.join(out))
    _ffi.check(L.cfm_tc_debug_buffer(None), 'dbg')
