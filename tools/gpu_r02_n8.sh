#!/bin/bash
# round-2 8-GPU visit: the driver's scaling command (bench.py --gpus 8 --steps 20 --warmup 5) with every config, the cross-GPU gather check,
# and the multicast transport for comparison
TAG=${1:-r02n8}
OUT=gpurun_out/$TAG
mkdir -p $OUT
N=${2:-8}
nvidia-smi topo -m > $OUT/topo.txt 2>&1
echo "== multi_gpu_check"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577 tests/multi_gpu_check.py 100000 > $OUT/multi_gpu_check.txt 2>&1; echo "rc=$?"; grep "multi-GPU\|rror" $OUT/multi_gpu_check.txt | head -10
echo "== bench n$N"; timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29578 bench.py --gpus $N --steps 20 --warmup 5 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err; echo "rc=$?"; head -c 3000 $OUT/bench_n$N.json; echo; tail -3 $OUT/bench_n$N.err
echo "== bench n$N multicast"; ARKMPC_GATHER=multicast timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29579 bench.py --gpus $N --steps 20 --warmup 5 --configs 4 --e2e-steps 0 > $OUT/bench_n${N}_multicast.json 2>> $OUT/bench_n$N.err; echo "rc=$?"; python -c "
import json;d=json.load(open('$OUT/bench_n${N}_multicast.json'));print(d.get('open_gather'));print(d['configs'][4].get('open_gather') if d['configs'][4] else None)"
