#!/bin/bash
echo "== stream tests"; timeout 600 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider -k "coupling_stream" 2>&1 | tail -3
echo "== stream diag"; timeout 300 python scripts/stream_diag.py 2>&1 | tail -9
