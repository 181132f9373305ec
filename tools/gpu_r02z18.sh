#!/bin/bash
# round-2 GPU visit Z18 (1 GPU): host path against the staging chunk size (ARKMPC_CHUNK_LOG2)
OUT=gpurun_out/r02z18; mkdir -p $OUT
for c in 17 18 19 20; do
  ARKMPC_CHUNK_LOG2=$c timeout 300 python bench.py --steps 20 --warmup 5 --configs none --no-cpu-baseline 2>> $OUT/bench.err | python -c "import json,sys;d=json.loads(sys.stdin.read());e=d['e2e'];print('chunk 2^$c: AoS', round(e['ms_per_step'],2), 'ms', round(e['value']/1e6,1), 'M; share planes', round(e['share_plane_operands']['ms_per_step'],2), 'ms', round(e['share_plane_operands']['value']/1e6,1), 'M')"
done
