#!/bin/bash
# round-2 session h: full suite + smoke, then the bench five times in a row (step-time outliers?), MLP timeline
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
echo "== full gpu suite" | tee gpurun_out/h_p1.log
timeout 1500 python -m pytest tests -q -m gpu --timeout 400 >> gpurun_out/h_p1.log 2>&1
echo "rc=$?" >> gpurun_out/h_p1.log; tail -6 gpurun_out/h_p1.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/h_smoke.log 2>&1; echo "smoke rc=$?"
timeout 300 python scripts/mlp_timeline.py > gpurun_out/h_timeline.log 2>&1; grep -v "tile . acc\|tile . drained" gpurun_out/h_timeline.log
for k in 1 2 3 4 5 6; do
  timeout 300 python bench.py --steps 20 --warmup 3 --no-ode --no-cpu-baseline --no-extra > gpurun_out/h_run$k.json 2> gpurun_out/h_run$k.err
  python -c "import json;d=json.load(open('gpurun_out/h_run$k.json'));print('run$k value',round(d['value'],2),'step_ms',{k:(round(v,3) if isinstance(v,float) else v) for k,v in d['step_ms'].items() if k!='note'},'stage',{k:round(v,3) for k,v in d['roofline']['stage_ms'].items()},'e2e',round(d['e2e']['value'],1),round(d['e2e']['blocking_call_value'],1))"
done
echo "== full bench"
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/h_bench.json 2> gpurun_out/h_bench.err; echo "bench rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/h_bench.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'blocking',d['e2e']['blocking_call_value'], d['step_ms'])
print('stage', d['roofline']['stage_ms'])
print('ode',d['ode']['value'],d['ode']['ms_per_trajectory'],'fwd_ms',d['ode']['mlp_forward_ms'])
print('c1',d['c1_coupling']['ms_per_coupling'],'c4',d['c4']['ms_per_shard_coupling'])"
