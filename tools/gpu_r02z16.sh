#!/bin/bash
# round-2 GPU visit Z16 (1 GPU): run-to-run spread of K2 with evict-first operand loads, with and without evict-last mask stores
OUT=gpurun_out/r02z16; mkdir -p $OUT
for k in 0 1; do for r in 1 2 3 4 5 6; do
  ARKMPC_L2_KEEP=$k timeout 300 python bench.py --steps 20 --warmup 5 --configs none --e2e-steps 0 --no-cpu-baseline 2>> $OUT/bench.err | python -c "import json,sys;d=json.loads(sys.stdin.read());print('keep $k run $r', round(d['value']/1e9,3), round(d['roofline']['kernel_us'],2), round(d['roofline']['frac'],4))"
done; done
for k in 0 1; do ARKMPC_L2_KEEP=$k timeout 300 python bench.py --steps 100 --log2-batch 22 --configs none --e2e-steps 0 --no-cpu-baseline 2>> $OUT/bench.err | python -c "import json,sys;d=json.loads(sys.stdin.read());print('2^22 keep $k', round(d['value']/1e9,3), round(d['roofline']['frac'],4))"; done
