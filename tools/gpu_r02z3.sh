#!/bin/bash
# round-2 GPU visit Z3 (1 GPU): final check — inversion with the single-launch path for <= 131072 elements, full GPU suite, smoke, bench
TAG=${1:-r02z3}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== bench_extra"; timeout 600 python tools/bench_extra.py > $OUT/bench_extra.txt 2>&1; grep -E "^---|inverse|fft" $OUT/bench_extra.txt
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "== bench"; timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/bench.json 2>> $OUT/bench.err; echo "bench rc=$?"; python -c "
import json;d=json.load(open('$OUT/bench.json'));print(d['value'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['share_plane_operands']['value']); print([(c or {}).get('value') for c in d['configs']]); print(d['configs'][2].get('bn254_g1')); c=d['configs'][0]; print(c['value'], c['cpu_baseline']['value'])"
tail -3 $OUT/bench.err
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 | tail -1 | head -c 300; echo
