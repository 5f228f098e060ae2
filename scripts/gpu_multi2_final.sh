#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l
echo "== nccl test"; timeout 600 python -m pytest tests/test_multi_gpu.py -q -m gpu --timeout 500 -p no:cacheprovider 2>&1 | tail -2 | cut -c1-300
echo "== bench N=2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "rc=$?"; tail -2 gpurun_out/bench_n2.err | cut -c1-200
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_n2.json').read().strip().splitlines()[-1])
print(2, 'couplings/s', round(d['value'], 2), 'ms/step', round(d['ms_per_step'], 3), 'e2e', d['e2e']['value'], d['e2e']['blocking_call_value'], 'ode', round(d.get('ode', {}).get('value', 0)), 'ode_c1', round(d['ode_c1']['value']), 'launches', d['gpu_launches'], d['clocks'])
PY
