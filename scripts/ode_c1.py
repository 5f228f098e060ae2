"""BASELINE config 1 sampling (2-D tutorial): MLP(dim=2, w=64, time_varying), B=1024, dopri5 1e-4,
t_span = linspace(0, 1, 100) -- launch/sync-bound regime."""
import sys, time, torch
sys.path.insert(0, '.')
import cfm_b200
from oracle import vector_field as vf
dev = torch.device('cuda:0')
torch.manual_seed(0)
mlp = cfm_b200.MLP(dim=2, w=64, time_varying=True).to(dev)
x = torch.randn(1024, 2, device=dev)
span = torch.linspace(0, 1, 100)
for fused in (False, True):
    node = cfm_b200.NeuralODE(cfm_b200.torch_wrapper(mlp), solver="dopri5", atol=1e-4, rtol=1e-4)
    node.use_fused_small = fused
    for _ in range(3): out = node.trajectory(x, span)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10): out = node.trajectory(x, span)
    torch.cuda.synchronize()
    print('fused', fused, 'ms per trajectory', (time.perf_counter() - t0) * 100, node.stats)
ref_m = vf.make_mlp(2, w=64, time_varying=True).to(dev)
ref_m.load_state_dict(mlp.state_dict())
f = lambda t, z: vf.wrapped_forward(ref_m, t, z)
with torch.no_grad():
    for _ in range(2): ref = vf.dopri5_trajectory(f, x, span.to(dev))
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3): ref = vf.dopri5_trajectory(f, x, span.to(dev))
    torch.cuda.synchronize()
print('torch eager ms per trajectory', (time.perf_counter() - t0) / 3 * 1e3)
ref_t = ref[0] if isinstance(ref, tuple) else ref
print('max abs diff vs eager driver', float((out - ref_t.reshape(out.shape)).abs().max()))
