#!/bin/bash
# round-2 GPU visit N (1 GPU): safegcd top + in-thread binary product tree for batch inversion; unrolled Keccak in the C++ host mirror
TAG=${1:-r02n}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest (fr + host_cpp + offline)"; timeout 1500 python -m pytest tests/test_gpu_fr.py tests/test_host_cpp.py tests/test_gpu_offline.py -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log
echo "== bench_extra"; timeout 600 python tools/bench_extra.py > $OUT/bench_extra.txt 2>&1; grep -E "^---|inverse|fft" $OUT/bench_extra.txt
echo "== config0 (C++ host)"; timeout 300 tools/host_bench/bench_config0 1024 30 1; timeout 300 tools/host_bench/bench_config0 1024 30 0; timeout 300 tools/host_bench/bench_config0 65536 10 1
echo "== ncu inverse"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"fr_inv_" -c 14 -o $OUT/prof_inv -f \
  python tools/bench_ntt_once.py > $OUT/ncu_full_inv.log 2>&1; echo "ncu rc=$?"
