#!/bin/bash
# round-2 first GPU session: new kernels first (each phase under its own timeout), then the full suite, then bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/a_smi.txt 2>&1
echo "== phase 1: round-2 tests without the MLP kernels" | tee gpurun_out/a_p1.log
timeout 900 python -m pytest tests/test_gpu_round2.py -q -m gpu --timeout 400 -k "not mlp and not dopri5" >> gpurun_out/a_p1.log 2>&1
echo "rc=$?" >> gpurun_out/a_p1.log; tail -25 gpurun_out/a_p1.log
echo "== phase 2: MLP tensor-core kernels" | tee gpurun_out/a_p2.log
timeout 600 python -m pytest tests/test_gpu_round2.py -q -m gpu --timeout 200 -k "mlp or dopri5" >> gpurun_out/a_p2.log 2>&1
echo "rc=$?" >> gpurun_out/a_p2.log; tail -25 gpurun_out/a_p2.log
echo "== phase 3: full gpu suite" | tee gpurun_out/a_p3.log
timeout 1500 python -m pytest tests -q -m gpu --timeout 400 --deselect tests/test_gpu_round2.py >> gpurun_out/a_p3.log 2>&1
echo "rc=$?" >> gpurun_out/a_p3.log; tail -25 gpurun_out/a_p3.log
echo "== phase 4: smoke + bench" | tee gpurun_out/a_p4.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" >> gpurun_out/a_p4.log 2>&1
echo "smoke rc=$?" >> gpurun_out/a_p4.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err
echo "bench rc=$?" >> gpurun_out/a_p4.log
tail -5 gpurun_out/a_p4.log; head -c 3000 gpurun_out/a_bench.json; tail -5 gpurun_out/a_bench.err
