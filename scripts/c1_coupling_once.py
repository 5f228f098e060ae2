import sys, torch, numpy as np
sys.path.insert(0, '.')
import cfm_b200
dev = torch.device('cuda:0')
g = torch.Generator().manual_seed(256)
x0, x1 = torch.randn(256, 2, generator=g).to(dev), torch.randn(256, 2, generator=g).to(dev)
fm = cfm_b200.ExactOptimalTransportConditionalFlowMatcher(sigma=0.0)
for _ in range(4): out = fm.sample_location_and_conditional_flow(x0, x1)
torch.cuda.synchronize()
print(fm.ot_sampler.last_info)
