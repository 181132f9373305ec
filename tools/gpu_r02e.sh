#!/bin/bash
# round-2 GPU visit E (1 GPU): point kernels with split two-pass multiplication + L1-cached / prefetched scratch tables
TAG=${1:-r02e}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu (curve, fabric, golden)"; timeout 1500 python -m pytest tests -m gpu -x -q -k "curve or fabric or golden or host or offline" > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 $OUT/pytest_gpu.log
echo "== bench_points 17"; timeout 600 python tools/bench_points.py 17 > $OUT/bench_points_17.txt 2>&1; cat $OUT/bench_points_17.txt
echo "== bench point_mul"; timeout 600 python bench.py --workload point_mul > $OUT/bench_point_mul.json 2> $OUT/bench.err; cat $OUT/bench_point_mul.json
timeout 600 python bench.py --workload point_mul --field bn254_fr > $OUT/bench_point_mul_bn254.json 2>> $OUT/bench.err; cat $OUT/bench_point_mul_bn254.json
echo "== ncu pt recombine"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pt_beaver_recombine -c 2 -o $OUT/prof_pt_recombine -f \
  python tools/bench_points.py 17 > $OUT/ncu_full_pt.log 2>&1; echo "ncu full pt rc=$?"
tail -3 $OUT/bench.err
