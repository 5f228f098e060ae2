#!/bin/bash
mkdir -p gpurun_out
echo "== tc timeline"; timeout 300 python scripts/tc_timeline.py > gpurun_out/tc_timeline.log 2>&1; echo "rc=$?"; cat gpurun_out/tc_timeline.log | tail -8
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu.log | cut -c1-300
echo "== bench"; timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err; cut -c1-2200 gpurun_out/bench.json
ls -la gpurun_out | head -30
