#!/bin/bash
# round-2 8-GPU visit X: the driver's scaling command on the final library (cached allocator under the gathered planes), and the gather check
TAG=${1:-r02x}
OUT=gpurun_out/$TAG
mkdir -p $OUT
N=${2:-8}
echo "== multi_gpu_check"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577 tests/multi_gpu_check.py 100000 > $OUT/multi_gpu_check.txt 2>&1; echo "rc=$?"; grep "multi-GPU\|rror" $OUT/multi_gpu_check.txt | head -3
echo "== bench n$N"; timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29578 bench.py --gpus $N --steps 20 --warmup 5 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err; echo "rc=$?"; python -c "
import json;d=json.load(open('$OUT/bench_n$N.json'));print(d['value'], d.get('value_with_open_gather'), d['e2e']['value'], d.get('open_gather')); print([ (c or {}).get('value') for c in d['configs']])"; tail -3 $OUT/bench_n$N.err
echo "== reference arm n$N"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29579 bench.py --impl reference --gpus $N --steps 3 --warmup 1 2>/dev/null | tail -1 | head -c 400; echo
