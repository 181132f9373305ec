#!/bin/bash
# round-2 GPU visit D (1 GPU): parity suite with the scratch-table point kernels and the hint test, point timings, ncu of the point recombine
TAG=${1:-r02d}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 $OUT/pytest_gpu.log
echo "== bench_points 17"; timeout 600 python tools/bench_points.py 17 > $OUT/bench_points_17.txt 2>&1; cat $OUT/bench_points_17.txt
echo "== bench point_mul"; timeout 600 python bench.py --workload point_mul > $OUT/bench_point_mul.json 2> $OUT/bench.err; cat $OUT/bench_point_mul.json
timeout 600 python bench.py --workload point_mul --field bn254_fr > $OUT/bench_point_mul_bn254.json 2>> $OUT/bench.err; cat $OUT/bench_point_mul_bn254.json
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/bench.json 2>> $OUT/bench.err; echo "rc=$?"; head -c 1500 $OUT/bench.json; echo
echo "== bench no-hint"; timeout 300 python bench.py --steps 200 --no-hint --configs none --e2e-steps 0 --no-cpu-baseline > $OUT/bench_nohint.json 2>> $OUT/bench.err; head -c 300 $OUT/bench_nohint.json; echo
echo "== ncu pt recombine"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pt_beaver_recombine -c 2 -o $OUT/prof_pt_recombine -f \
  python tools/bench_points.py 17 > $OUT/ncu_full_pt.log 2>&1; echo "ncu full pt rc=$?"
echo "== ncu recombine"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:beaver_recombine_kernel -s 4 -c 2 -o $OUT/prof_recombine -f \
  python bench.py --steps 2 --warmup 3 --e2e-steps 0 --no-cpu-baseline --configs none > $OUT/ncu_full_bench.log 2>&1; echo "ncu full rc=$?"
echo "== ncu launches"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
  python bench.py --steps 2 --warmup 3 --e2e-steps 0 --no-cpu-baseline --configs none > $OUT/ncu_launch_bench.log 2>&1; echo "ncu launches rc=$?"
ls -la $OUT
tail -3 $OUT/bench.err
