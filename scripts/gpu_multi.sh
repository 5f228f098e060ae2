#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -4
echo "== nccl test"; timeout 600 python -m pytest tests/test_multi_gpu.py -q -m gpu --timeout 500 -p no:cacheprovider > gpurun_out/pytest_multi.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/pytest_multi.log | cut -c1-300
echo "== bench N=1"; timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "rc=$?"
echo "== bench N=2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "rc=$?"; tail -3 gpurun_out/bench_n2.err
python - <<'PY'
import json
for n in (1, 2):
    try:
        d = json.loads(open(f'gpurun_out/bench_n{n}.json').read().strip().splitlines()[-1])
        print(n, 'couplings/s', round(d['value'], 2), 'ms/step', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value'], 2), 'ode', round(d.get('ode', {}).get('value', 0)), 'launches', d['gpu_launches'])
    except Exception as e:
        print(n, 'ERR', e)
PY
