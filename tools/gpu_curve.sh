#!/bin/bash
# GPU visit for the point gates: parity tests (scalar + curve), smoke, quick timing of the point kernels.
TAG=${1:-r01c}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/smi.txt 2>&1
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 $OUT/pytest_gpu.log
echo "== smoke"; timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $OUT/smoke.log
echo "== bench (short)"; timeout 600 python bench.py --steps 200 --e2e-steps 2 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cat $OUT/bench.json; tail -5 $OUT/bench.err
if [ -f tools/bench_points.py ]; then echo "== bench_points"; timeout 900 python tools/bench_points.py > $OUT/bench_points.txt 2>&1; echo "rc=$?"; tail -30 $OUT/bench_points.txt; fi
