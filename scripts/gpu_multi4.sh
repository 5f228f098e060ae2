#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | cut -c1-300
echo "== pytest gpu (incl. multi-gpu)"; timeout 1500 python -m pytest tests -m gpu -q --timeout 400 -p no:cacheprovider 2>&1 | tail -3
for n in 1 4; do
echo "== bench N=$n"
if [ $n = 1 ]; then timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err; fi
echo "rc=$?"; tail -2 gpurun_out/bench_n$n.err | cut -c1-200
done
echo "== reference arm under torchrun"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29578 bench.py --impl reference --gpus 4 --steps 3 --warmup 1 2>/dev/null | tail -1 | cut -c1-300
python - <<'PY'
import json
for n in (1, 4):
    try:
        d = json.loads(open(f'gpurun_out/bench_n{n}.json').read().strip().splitlines()[-1])
        print(n, 'couplings/s', round(d['value'], 2), 'ms/step', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value'], 2), 'ode', round(d.get('ode', {}).get('value', 0)), 'clocks', d['clocks'])
    except Exception as e:
        print(n, 'ERR', e)
PY
