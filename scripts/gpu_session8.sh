#!/bin/bash
mkdir -p gpurun_out
echo "== ncu full sinkhorn_v2"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:sinkhorn_v2 -s 3 -c 1 -o gpurun_out/prof_sinkhorn_v2 python bench.py --steps 1 --warmup 3 --no-ode --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
echo "== ncu full sqdist"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel.*SqDist -s 3 -c 1 -o gpurun_out/prof_sqdist_tc python bench.py --steps 1 --warmup 3 --no-ode --no-cpu-baseline > gpurun_out/ncu_full2.log 2>&1; echo "ncu full rc=$?"
echo "== ncu launches"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
echo "== exact timing"; timeout 300 python - > gpurun_out/exact_timing.log 2>&1 <<'PY'
import sys, time, torch, numpy as np
sys.path.insert(0, '.')
import cfm_b200
from oracle import coupling as oc
dev = torch.device('cuda:0')
for n, d in ((128, 8), (256, 2), (512, 16), (1024, 32), (2048, 64), (4096, 64)):
    g = torch.Generator().manual_seed(n)
    x0, x1 = torch.randn(n, d, generator=g), torch.randn(n, d, generator=g)
    s = cfm_b200.OTPlanSampler('exact', warn=False)
    a, b = x0.to(dev), x1.to(dev)
    for _ in range(2): s.sample_plan(a, b)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    reps = 20 if n <= 1024 else 3
    for _ in range(reps): s.sample_plan(a, b)
    torch.cuda.synchronize(); gpu = (time.perf_counter() - t0) / reps
    t0 = time.perf_counter()
    for _ in range(3): oc.sample_plan(x0, x1, 'exact')
    cpu = (time.perf_counter() - t0) / 3
    sig = s.get_map(a, b).argmax(1)
    ok = np.array_equal(sig, oc.assignment(oc.cost_matrix(x0, x1)))
    print(f"n={n} d={d} gpu {gpu*1e3:.3f} ms/coupling  cpu-oracle {cpu*1e3:.3f} ms  sigma_exact={ok} info={s.last_info}")
PY
cat gpurun_out/exact_timing.log
ls -la gpurun_out | head
