#!/bin/bash
# round-2 GPU visit Y (1 GPU): host path with x, y as share planes (arkmpc_fr_batch_mul_begin_host_shares)
TAG=${1:-r02y}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest host path"; timeout 900 python -m pytest tests/test_gpu_fr.py -x -q -m gpu -k "host_buffer" 2>&1 | tail -2
echo "== bench"; timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --configs none --no-cpu-baseline > $OUT/bench.json 2>> $OUT/bench.err; echo "bench rc=$?"; python -c "
import json;d=json.load(open('$OUT/bench.json'));print(d['value'], d['roofline']['frac']); print(json.dumps(d['e2e'], indent=1))"
tail -3 $OUT/bench.err
