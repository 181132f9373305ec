#!/bin/bash
for m in persistent full; do
  echo "=== ARKMPC_GRID=$m"
  ARKMPC_GRID=$m python tools/bench_extra.py 2>&1 | grep -A8 "2^20" | head -9
  ARKMPC_GRID=$m python bench.py --steps 500 --e2e-steps 0 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench value', d['value']/1e9, 'G ms/step', d['ms_per_step'], 'K2 us', d['roofline']['kernel_us'], 'step_hbm', d['step_hbm_gbs'])"
done
