#!/bin/bash
# round-2 GPU visit U (1 GPU): NTT block shapes (finer blocks stagger the load / barrier phases under the multiplier roof)
TAG=${1:-r02u}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for tt in 256 128; do for st in 512 256 128; do
  echo "== tile $tt strided $st"; ARKMPC_NTT_TILE_THREADS=$tt ARKMPC_NTT_STRIDE_THREADS=$st timeout 600 python tools/bench_extra.py 2>&1 | grep -E "^---|fft \(one" | paste - - | awk '{print $4, $(NF-12), $(NF-11)}' | tr '\n' ';'; echo
done; done
echo "== parity 128/256"; ARKMPC_NTT_TILE_THREADS=128 ARKMPC_NTT_STRIDE_THREADS=256 timeout 600 python -m pytest tests/test_gpu_ntt.py -x -q -m gpu 2>&1 | tail -2
echo "== bn254 point, auto block"; timeout 300 python tools/bench_pt_bn_once.py 20; timeout 300 python tools/bench_pt_bn_once.py 17
