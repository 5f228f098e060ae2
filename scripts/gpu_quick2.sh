#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu.log | cut -c1-300
for cfg in 1 2; do
echo "== bench CFM_SK_CONFIG=$cfg"; CFM_SK_CONFIG=$cfg timeout 600 python bench.py --steps 20 --warmup 3 --no-ode --no-cpu-baseline > gpurun_out/bench_cfg$cfg.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/bench_cfg$cfg.json')); print(d['value'], d['roofline']['stage_ms'], d['parity'])"
done
