#!/bin/bash
# round-2 GPU visit Z20 (1 GPU): inversion and NTT kernel chains launched with programmatic stream serialization
OUT=gpurun_out/r02z20; mkdir -p $OUT
echo "== pytest ntt + offline + fabric"; timeout 900 python -m pytest tests/test_gpu_ntt.py tests/test_gpu_offline.py tests/test_gpu_fabric.py -x -q -m gpu 2>&1 | tail -2
for p in 1 0; do echo "== ARKMPC_PDL=$p"; ARKMPC_PDL=$p timeout 600 python tools/bench_extra.py 2>&1 | grep -E "^---|inverse|fft"; done
echo "== memcheck"; timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_ntt.py -x -q -m gpu -k "inverse or fft" > $OUT/memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 $OUT/memcheck.log
