#!/bin/bash
# round-2 session e (2 GPUs): NCCL tests (sharded coupling, lock-step ODE) + bench under torchrun at N=2
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi -L
echo "== multi-gpu tests" | tee gpurun_out/e_p1.log
timeout 900 python -m pytest tests/test_multi_gpu.py -q -m gpu --timeout 600 >> gpurun_out/e_p1.log 2>&1
echo "rc=$?" >> gpurun_out/e_p1.log; tail -15 gpurun_out/e_p1.log
echo "== bench N=2"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/e_bench2.out 2> gpurun_out/e_bench2.err; echo "rc=$?"
grep "^{" gpurun_out/e_bench2.out > gpurun_out/e_bench2.json
python -c "
import json;d=json.load(open('gpurun_out/e_bench2.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'], d['step_ms'])
print('ode',d['ode']['value'],d['ode']['ms_per_trajectory'])
print('ode_strong',json.dumps(d.get('ode_strong'))[:800])
print('c4',d['c4']['ms_per_shard_coupling'], d['c4']['shard_couplings_per_s'])"
tail -5 gpurun_out/e_bench2.err
