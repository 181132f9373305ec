#!/bin/bash
# round-2 GPU visit B (1 GPU): parity suite incl. thread-safety / NCCL-ABI tests, new bench.py with configs, K2 timeline, zero-copy probe
TAG=${1:-r02b}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > $OUT/smi.txt 2>&1
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 $OUT/pytest_gpu.log
echo "== zc"; timeout 120 tools/_zc > $OUT/zc.txt 2>&1; cat $OUT/zc.txt
echo "== k2t"; timeout 120 tools/_k2t > $OUT/k2t.txt 2>&1; grep -v "^   \[" $OUT/k2t.txt
echo "== bench (zc)"; timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cat $OUT/bench.json; tail -5 $OUT/bench.err
echo "== bench e2e flat"; ARKMPC_XY=flat timeout 300 python bench.py --steps 200 --configs none --no-cpu-baseline > $OUT/bench_flat.json 2>> $OUT/bench.err; python -c "import json;d=json.load(open('$OUT/bench_flat.json'));print(d['e2e'])"
echo "== bench e2e 2d"; ARKMPC_XY=2d timeout 300 python bench.py --steps 200 --configs none --no-cpu-baseline > $OUT/bench_2d.json 2>> $OUT/bench.err; python -c "import json;d=json.load(open('$OUT/bench_2d.json'));print(d['e2e'])"
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; cat $OUT/bench_ref.json
