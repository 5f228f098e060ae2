#!/bin/bash
echo "== pytest exact"; timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider -k "exact or perm or trajectory" 2>&1 | tail -5 | cut -c1-300
echo "== timing (fast CTA kernel)"; timeout 300 python scripts/exact_timing.py 2>&1 | grep "n=" | head -12
echo "== timing (warp kernel up to 512)"; CFM_ASSIGN_WARP_MAX=512 timeout 300 python scripts/exact_timing.py --small 2>&1 | grep "n=" | head -3
