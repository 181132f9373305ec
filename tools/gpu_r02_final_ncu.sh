#!/bin/bash
# round-2: ncu evidence for the final build — launch list of the bench command and full captures of K2 / K1
TAG=${1:-r02zz}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== ncu launches"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
  python bench.py --steps 2 --warmup 3 --e2e-steps 0 --no-cpu-baseline --configs none > $OUT/ncu_launch_bench.log 2>&1; echo "ncu launches rc=$?"
echo "== ncu full K2 / K1"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:beaver_recombine_kernel -s 4 -c 2 -o $OUT/prof_recombine -f \
  python bench.py --steps 2 --warmup 3 --e2e-steps 0 --no-cpu-baseline --configs none > $OUT/ncu_full_bench.log 2>&1; echo "ncu full rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:beaver_mask -s 4 -c 2 -o $OUT/prof_mask -f \
  python bench.py --steps 2 --warmup 3 --e2e-steps 0 --no-cpu-baseline --configs none > $OUT/ncu_full_mask.log 2>&1; echo "ncu full mask rc=$?"
