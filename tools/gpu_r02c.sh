#!/bin/bash
# round-2 GPU visit C (2 GPUs): multicast probe, cross-GPU gather check (multicast / IPC / NCCL behind the ABI), 2-GPU bench
TAG=${1:-r02c}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1
echo "== mc_probe"; timeout 120 tools/mc_probe > $OUT/mc_probe.txt 2>&1; echo "rc=$?"; cat $OUT/mc_probe.txt
echo "== multi_gpu_check"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29577 tests/multi_gpu_check.py 200000 > $OUT/multi_gpu_check.txt 2>&1; echo "rc=$?"; grep -v "^\[W\|^W1\|^\*\*\*" $OUT/multi_gpu_check.txt | tail -15
echo "== bench n2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29578 bench.py --gpus 2 --steps 200 --warmup 5 --configs 4 > $OUT/bench_n2.json 2> $OUT/bench_n2.err; echo "rc=$?"; cat $OUT/bench_n2.json; tail -5 $OUT/bench_n2.err
echo "== bench n2 ipc"; ARKMPC_GATHER=ipc timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29579 bench.py --gpus 2 --steps 200 --warmup 5 --configs none --e2e-steps 0 > $OUT/bench_n2_ipc.json 2>> $OUT/bench_n2.err; echo "rc=$?"; python -c "import json;d=json.load(open('$OUT/bench_n2_ipc.json'));print(d.get('open_gather'), d.get('value_with_open_gather'))"
