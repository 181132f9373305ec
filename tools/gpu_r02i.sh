#!/bin/bash
# round-2 GPU visit I (1 GPU): product-tree batch inversion, shared-memory NTT twiddles, full parity suite, ncu of the NTT / inversion kernels
TAG=${1:-r02i}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 $OUT/pytest_gpu.log
echo "== bench_extra"; timeout 600 python tools/bench_extra.py > $OUT/bench_extra.txt 2>&1; cat $OUT/bench_extra.txt
echo "== ncu ntt / inverse"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"fr_ntt_tile|fr_ntt_strided|fr_inv_" -c 12 -o $OUT/prof_ntt -f \
  python tools/bench_ntt_once.py > $OUT/ncu_full_ntt.log 2>&1; echo "ncu full ntt rc=$?"
