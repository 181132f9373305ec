#!/bin/bash
# round-2 GPU visit O (1 GPU): where a 1024-gate iteration of the C++ host mirror spends its time
TAG=${1:-r02o}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for a in "1024 30 1" "65536 10 1"; do
  echo "== config0 $a"; ARKMPC_HOST_PROFILE=1 timeout 300 tools/host_bench/bench_config0 $a 2>&1 | tee -a $OUT/config0_profile.txt
done
nproc; lscpu | grep -i "model name"
