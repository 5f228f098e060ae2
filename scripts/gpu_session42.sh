#!/bin/bash
timeout 600 python scripts/small_sinkhorn_timing.py 2>&1 | grep "n=" 
