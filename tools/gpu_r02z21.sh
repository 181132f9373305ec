#!/bin/bash
# round-2 GPU visit Z21 (1 GPU): chains with the top inversion kernel launched plainly
for p in 1 0; do echo "== ARKMPC_PDL=$p"; ARKMPC_PDL=$p timeout 600 python tools/bench_extra.py 2>&1 | grep -E "^---|inverse"; done
echo "== pytest ntt"; timeout 900 python -m pytest tests/test_gpu_ntt.py -x -q -m gpu 2>&1 | tail -2
