#!/bin/bash
# round-2 third GPU session: tests, fused-MLP timeline (16 epilogue warps), Sinkhorn L2-prefetch sweep, cost tile A/B, bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
echo "== tests (round 2 file)" | tee gpurun_out/c_p1.log
timeout 900 python -m pytest tests/test_gpu_round2.py -q -m gpu --timeout 400 >> gpurun_out/c_p1.log 2>&1
echo "rc=$?" >> gpurun_out/c_p1.log; tail -8 gpurun_out/c_p1.log
echo "== timelines" | tee gpurun_out/c_timeline.log
timeout 300 python scripts/mlp_timeline.py >> gpurun_out/c_timeline.log 2>&1; grep -v "tile . acc\|tile . drained" gpurun_out/c_timeline.log; grep "L2 tile\|L4 tile 3" gpurun_out/c_timeline.log
for pf in 0 3 6 10 14; do
  echo "== CFM_SK_PF=$pf"
  CFM_SK_PF=$pf timeout 300 python bench.py --steps 10 --warmup 3 --no-ode --no-cpu-baseline --no-extra > gpurun_out/c_pf$pf.json 2> gpurun_out/c_pf$pf.err
  python -c "import json;d=json.load(open('gpurun_out/c_pf$pf.json'));print('pf=$pf value',round(d['value'],2),'stage_ms',{k:round(v,3) for k,v in d['roofline']['stage_ms'].items()})"
done
for l2 in 0.0 0.08 0.25; do
  echo "== CFM_SK_L2=$l2 (PF=6)"
  CFM_SK_L2=$l2 timeout 300 python bench.py --steps 10 --warmup 3 --no-ode --no-cpu-baseline --no-extra > gpurun_out/c_l2_$l2.json 2> gpurun_out/c_l2_$l2.err
  python -c "import json;d=json.load(open('gpurun_out/c_l2_$l2.json'));print('l2=$l2 value',round(d['value'],2),'stage_ms',{k:round(v,3) for k,v in d['roofline']['stage_ms'].items()})"
done
echo "== CFM_H3_TN=128"
CFM_H3_TN=128 timeout 300 python bench.py --steps 10 --warmup 3 --no-ode --no-cpu-baseline --no-extra > gpurun_out/c_tn128.json 2> gpurun_out/c_tn128.err
python -c "import json;d=json.load(open('gpurun_out/c_tn128.json'));print('tn128 value',round(d['value'],2),'stage_ms',{k:round(v,3) for k,v in d['roofline']['stage_ms'].items()})"
echo "== full bench"
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/c_bench.json 2> gpurun_out/c_bench.err; echo "bench rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/c_bench.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'blocking',d['e2e']['blocking_call_value'])
print('ode',d['ode']['value'],d['ode']['ms_per_trajectory'],'fwd_ms',d['ode']['mlp_forward_ms'])
print('c1',d['c1_coupling']['ms_per_coupling'],'c4',d['c4']['ms_per_shard_coupling'])"
