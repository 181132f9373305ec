#!/bin/bash
# round-2 GPU visit Z5 (1 GPU): Curve25519 point recombine, 512-thread blocks (128 registers) against 384 (168 registers)
for b in 512 384; do for l in 17 20; do ARKMPC_PT_ED_BLOCK=$b timeout 300 python tools/bench_pt_bn_once.py $l ed25519; done; done
echo "== parity with 384"; ARKMPC_PT_ED_BLOCK=384 timeout 900 python -m pytest tests/test_gpu_curve.py -x -q -m gpu -k "beaver" 2>&1 | tail -2
