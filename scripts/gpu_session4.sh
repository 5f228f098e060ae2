#!/bin/bash
mkdir -p gpurun_out
echo "== simt anomaly probe"; timeout 300 python - > gpurun_out/simt_probe.log 2>&1 <<'PY'
import torch, sys
sys.path.insert(0, '.')
from cfm_b200.optimal_transport import OTPlanSampler
dev = torch.device('cuda:0')
for trial in range(3):
    torch.manual_seed(trial)
    x0, x1 = torch.randn(512, 64, device='cuda'), torch.randn(768, 64, device='cuda')
    ref = (torch.cdist(x0.cpu().double(), x1.cpu().double()) ** 2)
    ref32 = torch.cdist(x0.cpu(), x1.cpu()) ** 2
    for algo in (1, 2, 1):
        s = OTPlanSampler('sinkhorn', cost_algo=algo)
        M, cmax, n0, n1 = s._cost(x0, x1, dev)
        torch.cuda.synchronize()
        e = (M[:, :n1].cpu().double() - ref).abs()
        bad = (e > 1e-3).nonzero()
        print('trial', trial, 'algo', algo, 'max err vs f64', e.max().item(), 'nbad', len(bad), bad[:6].tolist(),
              'cdist32 err', (ref32.double() - ref).abs().max().item())
PY
cat gpurun_out/simt_probe.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/pytest_gpu.log | cut -c1-300
echo "== bench"; timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -3 gpurun_out/bench.err; cut -c1-2200 gpurun_out/bench.json
echo "== ncu launches"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-ode --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
echo "== ncu full sinkhorn_v2"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:sinkhorn_v2 -s 3 -c 1 -o gpurun_out/prof_sinkhorn_v2 python bench.py --steps 1 --warmup 3 --no-ode --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
echo "== ncu full sqdist_tc"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:sqdist_tc -s 3 -c 1 -o gpurun_out/prof_sqdist_tc python bench.py --steps 1 --warmup 3 --no-ode --no-cpu-baseline > gpurun_out/ncu_full2.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out | head -30
