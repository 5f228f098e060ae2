#!/bin/bash
echo "== ode c1 dbg"; CFM_ODE_DBG=1 timeout 300 python scripts/ode_c1.py 2>&1 | grep -v "^$" | tail -4
echo "== pytest fused small"; timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider -k "fused_small or dopri5" 2>&1 | tail -5 | cut -c1-300
