"""C3 trajectory timing A/B: stage input formed inside the fused MLP kernel (one launch per NFE) vs the separate
stage-input kernel; graph on / off."""
import sys, time, torch
sys.path.insert(0, '.')
import cfm_b200
dev = torch.device('cuda:0')
torch.manual_seed(0)
mlp = cfm_b200.MLP(dim=784, w=256, time_varying=True).to(dev)
x = torch.randn(10000, 784, device=dev)
span = torch.linspace(0, 1, 2)
ref = None
import os
variants = [(False, True, 4), (True, True, 4), (True, True, 1), (True, False, 4), (False, True, 4), (True, True, 4)]
if len(sys.argv) > 1:  # e.g. "0" or "1" or "01": just these fuse settings, graph on
    variants = [(c == "1", True, 4) for c in sys.argv[1]]
print({k: v for k, v in os.environ.items() if k.startswith("CFM_")})
for fuse, graph, mb in variants:
    node = cfm_b200.NeuralODE(cfm_b200.torch_wrapper(mlp), solver="dopri5", atol=1e-4, rtol=1e-4)
    node.fuse_stage_input, node.use_cuda_graph, node.min_burst = fuse, graph, mb
    for _ in range(3):
        out = node.trajectory(x, span)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    for _ in range(10):
        out = node.trajectory(x, span)
    b.record(); torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / 10
    if ref is None:
        ref = out.clone()
    ms = a.elapsed_time(b) / 10
    print(f"fuse={fuse} graph={graph} min_burst={mb}: {ms:.3f} ms per trajectory ({10000 / ms / 1e3:.2f} M samples/s; wall {wall * 1e3:.3f}), "
          f"nfe {node.stats['nfe']}, bit-identical to first variant: {torch.equal(out, ref)}")
