#!/bin/bash
# round-2 GPU visit Z13 (1 GPU): safegcd in the base-field inversions (pt_normalize, fixed-base table build)
OUT=gpurun_out/r02z13; mkdir -p $OUT
echo "== pytest curve + golden + fabric + wire"; timeout 1500 python -m pytest tests/test_gpu_curve.py tests/test_golden_curve.py tests/test_gpu_fabric.py tests/test_wire.py -x -q -m gpu 2>&1 | tail -2
echo "== bench_points"; timeout 600 python tools/bench_points.py 17 2>&1 | tee $OUT/bench_points.txt | grep -E "normalize|recombine|two-party|generator"
echo "== memcheck"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_curve.py -x -q -m gpu -k "linear or sums" > $OUT/memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 $OUT/memcheck.log
