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
for overlap in (False, True):
    for graph in (True, False):
        node = cfm_b200.NeuralODE(cfm_b200.torch_wrapper(mlp), solver="dopri5", atol=1e-4, rtol=1e-4)
        node.overlap_stages, node.use_cuda_graph = overlap, graph
        for _ in range(3):
            out = node.trajectory(x, span)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            out = node.trajectory(x, span)
        b.record(); torch.cuda.synchronize()
        if ref is None:
            ref = out.clone()
        print(f"overlap={overlap} graph={graph}: {a.elapsed_time(b) / 10:.3f} ms per trajectory, stats {node.stats}, "
              f"bit-identical to first variant: {torch.equal(out, ref)}")
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
