#!/bin/bash
# round-2 GPU visit M (1 GPU): hint-level A/B, per-direction twiddle cache, configs[0] through the C++ host mirror
TAG=${1:-r02m}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log
for L in 0 1 2; do
  echo "== bench hint level $L"; timeout 300 python bench.py --steps 200 --hint-level $L --configs none --e2e-steps 0 --no-cpu-baseline > $OUT/bench_hint$L.json 2>> $OUT/bench.err
  python -c "import json;d=json.load(open('$OUT/bench_hint$L.json'));print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_us'], d['clocks']['sm_mhz'], d['clocks']['reasons'])"
done
for L in 1 2; do
  echo "== bench hint level $L (20 steps)"; timeout 300 python bench.py --steps 20 --warmup 5 --hint-level $L --configs none --e2e-steps 0 --no-cpu-baseline > $OUT/bench20_hint$L.json 2>> $OUT/bench.err
  python -c "import json;d=json.load(open('$OUT/bench20_hint$L.json'));print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_us'], d['clocks']['sm_mhz'], d['clocks']['reasons'])"
done
echo "== config0 (C++ host)"; timeout 300 tools/host_bench/bench_config0 1024 30 1; timeout 300 tools/host_bench/bench_config0 1024 30 0; timeout 300 tools/host_bench/bench_config0 65536 10 1
echo "== bench"; timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/bench.json 2>> $OUT/bench.err; echo "bench rc=$?"; python -c "
import json;d=json.load(open('$OUT/bench.json'));print(d['value'], d['roofline']['frac']); c=d['configs'][0]; print({k:c[k] for k in ('value','ms_per_iter','python_mirror','cpp_host_mirror','cpu_baseline')})"
tail -3 $OUT/bench.err
echo "== synccheck (phase-synchronised point kernels)"
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_gpu_curve.py -x -q -m gpu -k "point_beaver_mul_bit_exact or point_sums_and_msm or validation" > $OUT/synccheck.log 2>&1; echo "synccheck rc=$?"; tail -4 $OUT/synccheck.log
echo "== memcheck (point kernels)"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_curve.py -x -q -m gpu -k "point_beaver_mul_bit_exact or validation or scalar_mul" > $OUT/memcheck_curve.log 2>&1; echo "memcheck rc=$?"; tail -4 $OUT/memcheck_curve.log
