#!/bin/bash
for v in 0 1 2 3; do
  echo "=== ARKMPC_PT_RECOMBINE=$v"
  ARKMPC_PT_RECOMBINE=$v python tools/bench_points.py 18 2>&1 | grep -E "recombine|two-party"
done
