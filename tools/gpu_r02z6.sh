#!/bin/bash
# round-2 GPU visit Z6 (1 GPU): Curve25519 two-pass multiplications over four tables (P, 2^64 P, 2^128 P, 2^192 P)
TAG=${1:-r02z6}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest curve + fabric"; timeout 1500 python -m pytest tests/test_gpu_curve.py tests/test_golden_curve.py tests/test_gpu_fabric.py -x -q -m gpu 2>&1 | tail -2
for l in 17 20; do timeout 300 python tools/bench_pt_bn_once.py $l ed25519; timeout 300 python tools/bench_pt_bn_once.py $l; done
echo "== bench_points"; timeout 600 python tools/bench_points.py 17 2>&1 | tee $OUT/bench_points.txt | grep -E "pt_mul|recombine|two-party"
echo "== bench point_mul"; timeout 900 python bench.py --workload point_mul --steps 6 --warmup 3 > $OUT/bench_point_mul.json 2>> $OUT/bench.err; python -c "
import json;d=json.load(open('$OUT/bench_point_mul.json'));print(d['value'], d.get('bn254_g1'), d.get('parity'))"
tail -2 $OUT/bench.err
echo "== memcheck"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_curve.py -x -q -m gpu -k "matches_oracle or scalar_mul" > $OUT/memcheck_curve.log 2>&1; echo "memcheck rc=$?"; tail -3 $OUT/memcheck_curve.log
