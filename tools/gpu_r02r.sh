#!/bin/bash
# round-2 GPU visit R (1 GPU): lazy shared event sets in the memory cache, per-context staging rings, 64-thread inversion sweeps
TAG=${1:-r02r}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log
for a in "1024 30 1" "1024 30 0" "65536 10 1"; do
  echo "== config0 $a"; ARKMPC_HOST_PROFILE=1 timeout 300 tools/host_bench/bench_config0 $a 2>&1 | tee -a $OUT/config0_profile.txt
done
for b in 32 64 128; do echo "== inverse, block $b"; ARKMPC_INV_BLOCK=$b timeout 600 python tools/bench_extra.py 2>&1 | grep -E "^---|inverse"; done
echo "== thread test"; timeout 600 tests/host_cpp/test_threads 4 6 | tail -2
echo "== memcheck host mirror"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 tests/host_cpp/test_host > $OUT/memcheck_host.log 2>&1; echo "memcheck rc=$?"; tail -3 $OUT/memcheck_host.log
