#!/bin/bash
# round-2 GPU visit A: multicast probe, parity suite, bench with the constant-multiplier K2, TMA A/B, launch-shape harness
TAG=${1:-r02a}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/smi.txt 2>&1
nvidia-smi topo -m > $OUT/topo.txt 2>&1
echo "== mc_probe"; timeout 120 tools/mc_probe > $OUT/mc_probe.txt 2>&1; echo "rc=$?"; cat $OUT/mc_probe.txt
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_gpu.log
echo "== bench"; timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cat $OUT/bench.json; tail -3 $OUT/bench.err
echo "== bench tma"; ARKMPC_RECOMBINE=tma timeout 300 python bench.py --steps 500 --e2e-steps 0 --no-cpu-baseline > $OUT/bench_tma.json 2>> $OUT/bench.err; cat $OUT/bench_tma.json
echo "== bench c25519"; timeout 300 python bench.py --field curve25519_fr --steps 500 --e2e-steps 0 --no-cpu-baseline > $OUT/bench_c25519.json 2>> $OUT/bench.err; cat $OUT/bench_c25519.json
echo "== k2v"; timeout 120 tools/_k2v > $OUT/k2v.txt 2>&1; cat $OUT/k2v.txt
echo "== bench_extra"; timeout 300 python tools/bench_extra.py > $OUT/bench_extra.txt 2>&1; tail -30 $OUT/bench_extra.txt
