#!/bin/bash
# One GPU visit: parity tests, smoke, bench, pipe microbench, ncu launch list + full capture of the top kernel.
# Usage (under gpurun): bash tools/gpu_round.sh <tag>
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/smi.txt 2>&1
echo "== pytest -m gpu" ; timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_gpu.log
echo "== smoke" ; timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $OUT/smoke.log
echo "== pipe_bench"; timeout 120 ./tools/pipe_bench > $OUT/pipe_bench.txt 2>&1; tail -40 $OUT/pipe_bench.txt
echo "== bench"; timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cat $OUT/bench.json; tail -5 $OUT/bench.err
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; cat $OUT/bench_ref.json
echo "== ncu launches"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
  python bench.py --steps 2 --warmup 3 --e2e-steps 0 --no-cpu-baseline > $OUT/ncu_launch_bench.log 2>&1; echo "ncu launches rc=$?"
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:beaver_recombine -s 4 -c 2 -o $OUT/prof_recombine -f \
  python bench.py --steps 2 --warmup 3 --e2e-steps 0 --no-cpu-baseline > $OUT/ncu_full_bench.log 2>&1; echo "ncu full rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:beaver_mask -s 4 -c 2 -o $OUT/prof_mask -f \
  python bench.py --steps 2 --warmup 3 --e2e-steps 0 --no-cpu-baseline > $OUT/ncu_full_mask.log 2>&1; echo "ncu full mask rc=$?"
ls -la $OUT
