#!/bin/bash
# BN254 recombine with affine tables: 256- against 384-thread blocks at 2^19 / 2^20
for b in 256 384; do for l in 19 20; do ARKMPC_PT_BN_BLOCK=$b timeout 300 python tools/bench_pt_bn_once.py $l; done; done
