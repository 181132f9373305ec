#!/bin/bash
# round-2 GPU visit Z19 (1 GPU): host-path tests incl. the many-chunk case; default bench with the 2^19 chunk
OUT=gpurun_out/r02z19; mkdir -p $OUT
echo "== pytest host path"; timeout 900 python -m pytest tests/test_gpu_fr.py -x -q -m gpu -k "host_buffer" 2>&1 | tail -2
echo "== bench"; timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/bench.json 2>> $OUT/bench.err; echo "bench rc=$?"; python -c "
import json;d=json.load(open('$OUT/bench.json'));print(d['value'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['share_plane_operands']['value'], d['e2e']['share_plane_operands']['ms_per_step']); print([(c or {}).get('value') for c in d['configs']]); print(d['configs'][2].get('bn254_g1'))"
tail -2 $OUT/bench.err
