#!/bin/bash
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-200
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -2 gpurun_out/bench.err | cut -c1-200
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print('value', round(d['value'], 2), 'ms', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value'], 2), 'blocking', round(d['e2e']['blocking_call_value'], 2), 'ode', round(d['ode']['value']), 'ode_c1', round(d['ode_c1']['value']), 'cpu', d['cpu_baseline']['value'], 'clocks', d['clocks'])
PY
