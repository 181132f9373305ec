#!/bin/bash
# round-2 GPU visit V (1 GPU): Karatsuba field product in the NTT butterflies and the inversion sweeps, A/B against the interleaved product
TAG=${1:-r02v}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for m in cios kara; do
  echo "== ARKMPC_MUL=$m"; ARKMPC_MUL=$m timeout 600 python tools/bench_extra.py 2>&1 | grep -E "^---|inverse|fft" | tee $OUT/bench_extra_$m.txt
done
echo "== parity (kara)"; ARKMPC_MUL=kara timeout 900 python -m pytest tests/test_gpu_ntt.py -x -q -m gpu 2>&1 | tail -2
