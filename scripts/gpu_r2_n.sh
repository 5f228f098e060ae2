#!/bin/bash
# bench under torchrun at N GPUs (N = $1): weak-scaled C2, C4 shards, C5 sweep, ODE weak + strong (lock-step / independent)
N=${1:-4}
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/n${N}_bench.out 2> gpurun_out/n${N}_bench.err; echo "rc=$?"
grep "^{" gpurun_out/n${N}_bench.out > gpurun_out/n${N}_bench.json
python -c "
import json;d=json.load(open('gpurun_out/n${N}_bench.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'], d['step_ms'])
print('ode',d['ode']['value'],d['ode']['ms_per_trajectory'])
print('ode_strong',json.dumps(d.get('ode_strong'))[:900])
print('c4',d['c4']['ms_per_shard_coupling'], d['c4']['shard_couplings_per_s'])
print('c5',[(s['n_per_shard'],round(s['ms_per_shard_coupling'],3)) for s in d['c5']['sweep']])"
tail -3 gpurun_out/n${N}_bench.err
