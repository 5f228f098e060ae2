#!/bin/bash
for f in 0.1 0.15 0.25; do
CFM_SK_L2=$f timeout 300 python bench.py --steps 10 --warmup 3 --no-ode --no-cpu-baseline | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('L2frac $f', round(d['value'],1), d['roofline']['stage_ms']['solve'])"
done
for tn in 256 128; do
CFM_SK_L2=0.2 CFM_TC_TN=$tn timeout 300 python bench.py --steps 10 --warmup 3 --no-ode --no-cpu-baseline | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('TC_TN $tn', round(d['value'],1), d['roofline']['stage_ms'])"
done
