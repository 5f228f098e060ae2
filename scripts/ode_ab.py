"""C3 trajectory timing A/B: overlapped stage inputs on / off, graph on / off; plus three ways of timing one forward."""
import sys, time, torch
sys.path.insert(0, '.')
import cfm_b200
dev = torch.device('cuda:0')
torch.manual_seed(0)
mlp = cfm_b200.MLP(dim=784, w=256, time_varying=True).to(dev)
x = torch.randn(10000, 784, device=dev)
span = torch.linspace(0, 1, 2)
ref = None
variants = [(False, True, 1), (False, True, 4), (True, True, 1), (True, True, 4), (False, False, 4), (True, False, 4),
            (False, True, 1), (True, True, 4)]
for overlap, graph, mb in variants:
    node = cfm_b200.NeuralODE(cfm_b200.torch_wrapper(mlp), solver="dopri5", atol=1e-4, rtol=1e-4)
    node.overlap_stages, node.use_cuda_graph, node.min_burst = overlap, graph, mb
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
    print(f"overlap={overlap} graph={graph} min_burst={mb}: {a.elapsed_time(b) / 10:.3f} ms per trajectory (wall {wall * 1e3:.3f}), "
          f"nfe {node.stats['nfe']}, bit-identical to first variant: {torch.equal(out, ref)}")
# forward timing three ways
y = torch.empty_like(x)
t_dev = torch.full((1,), 0.5, device=dev)
with torch.no_grad():
    for _ in range(3):
        mlp.vector_field(t_dev, x, out=y)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(50):
        mlp.vector_field(t_dev, x, out=y)
    b.record(); torch.cuda.synchronize()
    print('eager loop, device t: %.1f us per forward' % (a.elapsed_time(b) * 20))
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        mlp.vector_field(t_dev, x, out=y)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    a.record()
    for _ in range(50):
        g.replay()
    b.record(); torch.cuda.synchronize()
    print('graph replay: %.1f us per forward' % (a.elapsed_time(b) * 20))
