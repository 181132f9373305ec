#!/bin/bash
# quick A/B: parity tests + bench (device-resident only) for the kernel variants
TAG=${1:-q}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_gpu.log
for f in bn254_fr curve25519_fr; do
  echo "== tma $f"; timeout 300 python bench.py --field $f --steps 500 --e2e-steps 0 --no-cpu-baseline > $OUT/bench_tma_$f.json 2>$OUT/err.log; python -c "import json;d=json.load(open('$OUT/bench_tma_$f.json'));print(d['value'], d['ms_per_step'], d['roofline']['kernel_us'], d['roofline']['frac'], d['clocks'])" || tail -5 $OUT/err.log
  echo "== ldg $f"; ARKMPC_RECOMBINE=ldg timeout 300 python bench.py --field $f --steps 500 --e2e-steps 0 --no-cpu-baseline > $OUT/bench_ldg_$f.json 2>$OUT/err.log; python -c "import json;d=json.load(open('$OUT/bench_ldg_$f.json'));print(d['value'], d['ms_per_step'], d['roofline']['kernel_us'], d['roofline']['frac'])" || tail -5 $OUT/err.log
done
