"""BASELINE config 4 shard (N=4096, d=512, eps=0.1 un-normalised: float64-potential mode): a few couplings, timed."""
import sys, torch
sys.path.insert(0, '.')
import cfm_b200
dev = torch.device('cuda:0')
g = torch.Generator().manual_seed(40)
x0, x1 = torch.randn(4096, 512, generator=g).to(dev), torch.randn(4096, 512, generator=g).to(dev)
fm = cfm_b200.SchrodingerBridgeConditionalFlowMatcher(sigma=0.05 ** 0.5, ot_method="sinkhorn")
sb = fm.ot_sampler
sb.num_iter_max, sb.stop_thr, sb.warn = 100, 0.0, False
n = int(sys.argv[1]) if len(sys.argv) > 1 else 5
for _ in range(2):
    sb.sample_plan(x0, x1)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(n):
    sb.sample_plan(x0, x1)
b.record(); torch.cuda.synchronize()
print(f"C4 shard coupling: {a.elapsed_time(b) / n:.3f} ms  info={sb.last_info}")
