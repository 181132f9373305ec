#!/bin/bash
# 2-GPU check of the final bench line (share-plane e2e pass under torchrun)
OUT=gpurun_out/r02s2; mkdir -p $OUT
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29578 bench.py --gpus 2 --steps 20 --warmup 5 > $OUT/bench_n2.json 2> $OUT/bench_n2.err; echo "rc=$?"; python -c "
import json;d=json.load(open('$OUT/bench_n2.json'));print(d['value'], d.get('value_with_open_gather'), d['e2e']['value'], d['e2e']['share_plane_operands']['value']); print([ (c or {}).get('value') for c in d['configs']])"; tail -3 $OUT/bench_n2.err
