#!/bin/bash
# round-2 GPU visit Q (1 GPU): inversion tree with stored inner products; default bench with the cached allocator in the host mirror
TAG=${1:-r02q}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest (fr + offline + fabric)"; timeout 1500 python -m pytest tests/test_gpu_fr.py tests/test_gpu_offline.py tests/test_gpu_fabric.py -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log
echo "== bench_extra"; timeout 600 python tools/bench_extra.py > $OUT/bench_extra.txt 2>&1; grep -E "^---|inverse|fft" $OUT/bench_extra.txt
echo "== ncu inverse"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"fr_inv_" -c 10 -o $OUT/prof_inv -f \
  python tools/bench_ntt_once.py > $OUT/ncu_full_inv.log 2>&1; echo "ncu rc=$?"
echo "== bench"; timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/bench.json 2>> $OUT/bench.err; echo "bench rc=$?"; python -c "
import json;d=json.load(open('$OUT/bench.json'));print(d['value'], d['roofline']['frac'], d['e2e']); c=d['configs'][0]; print({k:c[k] for k in ('value','ms_per_iter','python_mirror','cpp_host_mirror')})"
tail -3 $OUT/bench.err
