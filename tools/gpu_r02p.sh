#!/bin/bash
# round-2 GPU visit P (1 GPU): device-memory cache + staged uploads (C ABI), host mirror without redundant syncs
TAG=${1:-r02p}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_gpu.log
for a in "1024 30 1" "1024 30 0" "65536 10 1"; do
  echo "== config0 $a"; ARKMPC_HOST_PROFILE=1 timeout 300 tools/host_bench/bench_config0 $a 2>&1 | tee -a $OUT/config0_profile.txt
done
echo "== config0 with the cache off"; ARKMPC_ALLOC_CACHE_MB=0 timeout 300 tools/host_bench/bench_config0 1024 30 1
echo "== thread test"; timeout 600 tests/host_cpp/test_threads 4 6 | tail -2
echo "== memcheck host mirror"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 tests/host_cpp/test_host > $OUT/memcheck_host.log 2>&1; echo "memcheck rc=$?"; tail -3 $OUT/memcheck_host.log
echo "== racecheck-free ordering: initcheck"; timeout 900 compute-sanitizer --tool initcheck --error-exitcode 9 tools/host_bench/bench_config0 1024 3 1 > $OUT/initcheck.log 2>&1; echo "initcheck rc=$?"; tail -3 $OUT/initcheck.log
