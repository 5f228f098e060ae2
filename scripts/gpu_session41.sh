#!/bin/bash
echo "== stream diag"; timeout 300 python scripts/stream_diag.py 2>&1 | tail -12
echo "== tc timeline"; timeout 300 python scripts/tc_timeline.py 2>&1 | tail -4
