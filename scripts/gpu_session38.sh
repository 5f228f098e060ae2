#!/bin/bash
echo "== pytest exact"; timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider -k "exact or perm or trajectory or wasserstein" 2>&1 | tail -5 | cut -c1-300
echo "== timing (fast CTA kernel, column-reduction init)"; timeout 300 python scripts/exact_timing.py 2>&1 | grep "n=" | head -12
