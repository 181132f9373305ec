#!/bin/bash
# round-2 GPU visit Z12 (1 GPU): NTT tile kernel with the first double stage peeled (one multiplication instead of four) and the inverse scale folded in
OUT=gpurun_out/r02z12; mkdir -p $OUT
echo "== pytest ntt + fabric"; timeout 900 python -m pytest tests/test_gpu_ntt.py tests/test_gpu_fabric.py -x -q -m gpu 2>&1 | tail -2
echo "== bench_extra"; timeout 600 python tools/bench_extra.py 2>&1 | grep -E "^---|fft" | tee $OUT/bench_extra.txt
echo "== memcheck ntt"; timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_ntt.py -x -q -m gpu -k "fft" > $OUT/memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 $OUT/memcheck.log
