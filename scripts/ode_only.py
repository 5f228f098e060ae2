import sys, torch
sys.path.insert(0, '.')
import cfm_b200
dev = torch.device('cuda:0')
torch.manual_seed(0)
mlp = cfm_b200.MLP(dim=784, w=256, time_varying=True).to(dev)
node = cfm_b200.NeuralODE(cfm_b200.torch_wrapper(mlp), solver="dopri5", atol=1e-4, rtol=1e-4)
node.use_cuda_graph = '--eager' not in sys.argv
x = torch.randn(10000, 784, device=dev)
span = torch.linspace(0, 1, 2)
for _ in range(3): node.trajectory(x, span)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5): node.trajectory(x, span)
b.record(); torch.cuda.synchronize()
print('ms per trajectory', a.elapsed_time(b) / 5, node.stats)
