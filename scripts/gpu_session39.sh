#!/bin/bash
mkdir -p gpurun_out
for dbg in 1 5 1 5; do
CFM_SK_DBG=$dbg timeout 300 python bench.py --steps 20 --warmup 3 --no-ode --no-cpu-baseline > gpurun_out/b_$dbg.json 2>/dev/null
python - <<PY
import json
d=json.loads(open('gpurun_out/b_$dbg.json').read().strip().splitlines()[-1])
print('CFM_SK_DBG=$dbg', 'ms/step', round(d['ms_per_step'],3), 'solve', round(d['roofline']['stage_ms']['solve'],3), 'col err', d['parity']['col_marginal_max_rel_err'])
PY
done
