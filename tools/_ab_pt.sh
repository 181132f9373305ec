#!/bin/bash
for b in 2 3 4 0; do
  echo "=== ARKMPC_PT_BLOCKS=$b"
  ARKMPC_PT_BLOCKS=$b python tools/bench_points.py 18 2>&1 | grep -E "pt_mul \(var|recombine|two-party"
done
