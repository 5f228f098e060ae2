import sys, torch
sys.path.insert(0, '.')
import cfm_b200
dev = torch.device('cuda:0')
torch.manual_seed(0)
mlp = cfm_b200.MLP(dim=2, w=64, time_varying=True).to(dev)
x = torch.randn(1024, 2, device=dev)
span = torch.linspace(0, 1, 100)
node = cfm_b200.NeuralODE(cfm_b200.torch_wrapper(mlp), solver="dopri5", atol=1e-4, rtol=1e-4)
for _ in range(3): out = node.trajectory(x, span)
torch.cuda.synchronize()
print(node.stats)
