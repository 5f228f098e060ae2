#!/bin/bash
mkdir -p gpurun_out
echo "== pytest fused small"; timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider -k "fused_small or dopri5" 2>&1 | tail -25 | cut -c1-300
echo "== ode c1"; timeout 300 python scripts/ode_c1.py 2>&1 | tail -6
