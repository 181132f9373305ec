#!/bin/bash
for b in 0 3 4; do
  echo "=== ARKMPC_PT_MINB=$b"
  ARKMPC_PT_MINB=$b python tools/bench_points.py 18 2>&1 | grep -E "pt_mul \(var|recombine|two-party"
done
