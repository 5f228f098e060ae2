#!/bin/bash
# round-2 second GPU session: fixed tests, timelines, ncu captures of the fused MLP and the fp16x3 cost GEMM
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
echo "== tests (round 2 file)" | tee gpurun_out/b_p1.log
timeout 900 python -m pytest tests/test_gpu_round2.py -q -m gpu --timeout 400 -k "not c2_full" >> gpurun_out/b_p1.log 2>&1
echo "rc=$?" >> gpurun_out/b_p1.log; tail -12 gpurun_out/b_p1.log
echo "== timelines" | tee gpurun_out/b_timeline.log
timeout 300 python scripts/mlp_timeline.py >> gpurun_out/b_timeline.log 2>&1; cat gpurun_out/b_timeline.log
echo "== ncu fused mlp"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:mlp_fused --launch-skip 2 -c 1 -f -o gpurun_out/b_mlp_fused python scripts/mlp_once.py > gpurun_out/b_ncu1.log 2>&1; echo "rc=$?"
echo "== ncu cost gemm"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:gemm_h3 --launch-skip 2 -c 1 -f -o gpurun_out/b_sqdist_h3 python bench.py --steps 1 --warmup 3 --no-ode --no-cpu-baseline --no-extra > gpurun_out/b_ncu2.log 2>&1; echo "rc=$?"
ls -la gpurun_out/*.ncu-rep
