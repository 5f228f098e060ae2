#!/bin/bash
# round-2 session i: full suite after the ODE-driver / draw changes, smoke, bench x2, ODE-only timing
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
echo "== full gpu suite" | tee gpurun_out/i_p1.log
timeout 1500 python -m pytest tests -q -m gpu --timeout 400 >> gpurun_out/i_p1.log 2>&1
echo "rc=$?" >> gpurun_out/i_p1.log; tail -15 gpurun_out/i_p1.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/i_smoke.log 2>&1; echo "smoke rc=$?"
timeout 300 python scripts/ode_only.py > gpurun_out/i_ode.log 2>&1; tail -2 gpurun_out/i_ode.log
for k in 1 2; do
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/i_bench$k.json 2> gpurun_out/i_bench$k.err; echo "bench rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/i_bench$k.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'blocking',d['e2e']['blocking_call_value'], {k:(round(v,3) if isinstance(v,float) else v) for k,v in d['step_ms'].items() if k!='note'})
print('stage', d['roofline']['stage_ms'])
print('ode',d['ode']['value'],d['ode']['ms_per_trajectory'],'fwd_ms',d['ode']['mlp_forward_ms'])
print('c1',d['c1_coupling']['ms_per_coupling'],'c4',d['c4']['ms_per_shard_coupling'])"
done
