#!/bin/bash
# e2e (host-buffer path) vs staging chunk size; prints ms per 2^20-gate two-party step
for c in 13 14 15 16 17 18 20; do
  ARKMPC_CHUNK_LOG2=$c timeout 300 python bench.py --steps 20 --warmup 3 --e2e-steps 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('chunk 2^$c', 'e2e ms', round(d['e2e']['ms_per_step'],3), 'e2e mults/s', round(d['e2e']['value']/1e6,2),'M', 'h2d GB/s', round(d['e2e']['h2d_bytes_per_step']/d['e2e']['ms_per_step']/1e6,1))"
done
nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.width.current --format=csv
