#!/bin/bash
# Session 3: V2 Sinkhorn + tcgen05 cost (guarded by timeouts), full tests, bench, ncu evidence.
mkdir -p gpurun_out
echo "== tc cost quick"; timeout 120 python - > gpurun_out/tc_quick.log 2>&1 <<'PY'
import torch, sys
sys.path.insert(0, '.')
from cfm_b200.optimal_transport import OTPlanSampler
torch.manual_seed(0)
x0, x1 = torch.randn(512, 64, device='cuda'), torch.randn(768, 64, device='cuda')
dev = torch.device('cuda:0')
for algo in (1, 2):
    s = OTPlanSampler('exact', cost_algo=algo)
    M, cmax, n0, n1 = s._cost(x0, x1, dev)
    torch.cuda.synchronize()
    ref = torch.cdist(x0.cpu(), x1.cpu()) ** 2
    print('algo', algo, 'maxabs', (M[:, :n1].cpu() - ref).abs().max().item(), 'cmax', cmax.item(), ref.max().item())
PY
echo "tc rc=$?"; tail -5 gpurun_out/tc_quick.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/pytest_gpu.log | cut -c1-300
echo "== bench"; timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -3 gpurun_out/bench.err; cut -c1-2500 gpurun_out/bench.json
echo "== ncu launches"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-ode --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
echo "== ncu full sinkhorn"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:sinkhorn -s 3 -c 1 -o gpurun_out/prof_sinkhorn python bench.py --steps 1 --warmup 3 --no-ode --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out
