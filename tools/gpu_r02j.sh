#!/bin/bash
# round-2 GPU visit J (1 GPU): inversion tree with the branch-free top, radix-4 NTT tile + twiddle prefetch, wire format tests
TAG=${1:-r02j}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 $OUT/pytest_gpu.log
echo "== bench_extra"; timeout 600 python tools/bench_extra.py > $OUT/bench_extra.txt 2>&1; grep -E "^---|inverse|fft|fr_mul|mac_check|share_mul" $OUT/bench_extra.txt
echo "== ncu ntt / inverse"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"fr_ntt_tile|fr_ntt_strided|fr_inv_" -c 12 -o $OUT/prof_ntt -f \
  python tools/bench_ntt_once.py > $OUT/ncu_full_ntt.log 2>&1; echo "ncu full ntt rc=$?"
