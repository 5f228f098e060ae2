#!/bin/bash
# round-2 session d: MLP tests after the epilogue changes, timeline, step-time outlier experiment (clock sampler on/off), full suite + bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
echo "== full gpu suite" | tee gpurun_out/d_p1.log
timeout 1500 python -m pytest tests -q -m gpu --timeout 400 >> gpurun_out/d_p1.log 2>&1
echo "rc=$?" >> gpurun_out/d_p1.log; tail -6 gpurun_out/d_p1.log
echo "== timelines" | tee gpurun_out/d_timeline.log
timeout 300 python scripts/mlp_timeline.py >> gpurun_out/d_timeline.log 2>&1; grep -v "tile . acc\|tile . drained" gpurun_out/d_timeline.log; grep "L2 tile\|L4 tile 3" gpurun_out/d_timeline.log
for k in 1 2 3 4 5; do
  for mode in clk noclk; do
    if [ $mode = noclk ]; then export CFM_BENCH_NOCLK=1; else unset CFM_BENCH_NOCLK; fi
    timeout 300 python bench.py --steps 20 --warmup 3 --no-ode --no-cpu-baseline --no-extra > gpurun_out/d_$mode$k.json 2> gpurun_out/d_$mode$k.err
    python -c "import json;d=json.load(open('gpurun_out/d_$mode$k.json'));print('$mode$k value',round(d['value'],2),'step_ms',{k:round(v,3) for k,v in d['step_ms'].items() if k!='note'},'e2e',round(d['e2e']['value'],1),round(d['e2e']['blocking_call_value'],1))"
  done
done
unset CFM_BENCH_NOCLK
echo "== full bench"
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/d_bench.json 2> gpurun_out/d_bench.err; echo "bench rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/d_bench.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'blocking',d['e2e']['blocking_call_value'], d['step_ms'])
print('stage', d['roofline']['stage_ms'])
print('ode',d['ode']['value'],d['ode']['ms_per_trajectory'],'fwd_ms',d['ode']['mlp_forward_ms'])
print('c1',d['c1_coupling']['ms_per_coupling'],'c4',d['c4']['ms_per_shard_coupling'])"
