#!/bin/bash
# round-2 evidence pass: ncu launch list of the bench command + one full capture per dominant kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
echo "== timeline"; timeout 300 python scripts/mlp_timeline.py > gpurun_out/f_timeline.log 2>&1; grep -v "tile . acc\|tile . drained" gpurun_out/f_timeline.log; grep "L1 tile\|L3 tile 0\|L4 tile 3" gpurun_out/f_timeline.log
echo "== launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/f_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/f_ncu_bench.log 2>&1; echo "rc=$?"
echo "== full: sinkhorn_v2"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:sinkhorn_v2 -s 3 -c 1 -f -o gpurun_out/f_sinkhorn_v2 python bench.py --steps 1 --warmup 3 --no-ode --no-cpu-baseline --no-extra > gpurun_out/f_ncu1.log 2>&1; echo "rc=$?"
echo "== full: cost gemm"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_h3 -s 3 -c 1 -f -o gpurun_out/f_sqdist_h3 python bench.py --steps 1 --warmup 3 --no-ode --no-cpu-baseline --no-extra > gpurun_out/f_ncu2.log 2>&1; echo "rc=$?"
echo "== full: fused mlp"; timeout 600 ncu --set full --import-source on --clock-control none -k regex:mlp_fused --launch-skip 2 -c 1 -f -o gpurun_out/f_mlp_fused python scripts/mlp_once.py > gpurun_out/f_ncu3.log 2>&1; echo "rc=$?"
echo "== full: rk stage input"; timeout 600 ncu --set full --clock-control none -k regex:rk_stage_input --launch-skip 8 -c 1 -f -o gpurun_out/f_rk_stage python scripts/ode_only.py --eager > gpurun_out/f_ncu4.log 2>&1; echo "rc=$?"
ls -la gpurun_out/f_*.ncu-rep gpurun_out/f_launches.csv
