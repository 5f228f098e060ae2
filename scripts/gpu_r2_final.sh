#!/bin/bash
# round-2 final evidence pass: smoke, the whole GPU suite, the bench line, the ncu launch list of the bench command and
# full captures of the kernels that changed after scripts/gpu_r2_f.sh (stage-input kernel, float64-potential Sinkhorn)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-300
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/final_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/final_pytest.log | cut -c1-300
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/final_bench.err | cut -c1-200
python - <<'PY'
import json
d = json.loads(open('gpurun_out/final_bench.json').read().strip().splitlines()[-1])
print('value', round(d['value'], 2), 'ms', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value'], 2), 'blocking', round(d['e2e']['blocking_call_value'], 2))
print('ode', round(d['ode']['value']), d['ode']['ms_per_trajectory'], 'ode_c1', round(d['ode_c1']['value']))
print('c1', d['c1_coupling']['ms_per_coupling'], 'c4', d['c4']['ms_per_shard_coupling'], 'c5', [(s['n_per_shard'], round(s['ms_per_shard_coupling'], 3)) for s in d['c5']['sweep']])
print('roofline', {k: d['roofline'][k] for k in ('frac', 'frac_dram', 'achieved')}, 'stage_ms', d['roofline']['stage_ms'], 'clocks', d['clocks'])
PY
echo "== launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/final_ncu_bench.log 2>&1; echo "rc=$?"
echo "== full: rk stage input"; timeout 600 ncu --set full --clock-control none -k regex:rk_stage_input --launch-skip 8 -c 6 -f -o gpurun_out/final_rk_stage python scripts/ode_only.py --eager > gpurun_out/final_ncu4.log 2>&1; echo "rc=$?"
echo "== full: float64-potential sinkhorn (C4 shard)"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:sinkhorn_kernel -s 1 -c 1 -f -o gpurun_out/final_sinkhorn_c4 python scripts/c4_once.py 1 > gpurun_out/final_ncu5.log 2>&1; echo "rc=$?"
ls -la gpurun_out/final_*.ncu-rep gpurun_out/final_launches.csv
