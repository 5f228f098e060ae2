#!/bin/bash
# 2-GPU pass: NCCL tests + bench under torchrun (weak-scaled C2, C4 shards, C5 sweep, ODE weak + strong)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi -L | wc -l
echo "== nccl tests"; timeout 600 python -m pytest tests/test_multi_gpu.py -q -m gpu --timeout 500 -p no:cacheprovider 2>&1 | tail -2 | cut -c1-300
bash scripts/gpu_r2_n.sh 2
