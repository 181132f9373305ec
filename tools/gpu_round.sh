#!/bin/bash
# One GPU visit: parity tests, smoke, bench (+ reference arm), point timings, ncu launch list + full captures of the top kernels.
# Usage (under gpurun): bash tools/gpu_round.sh <tag>
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/smi.txt 2>&1
echo "== pytest -m gpu" ; timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_gpu.log
echo "== smoke" ; timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $OUT/smoke.log
echo "== bench"; timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cat $OUT/bench.json; tail -5 $OUT/bench.err
echo "== bench curve25519_fr"; timeout 600 python bench.py --field curve25519_fr --steps 500 --e2e-steps 0 --no-cpu-baseline > $OUT/bench_c25519.json 2>> $OUT/bench.err; cat $OUT/bench_c25519.json
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; cat $OUT/bench_ref.json
echo "== bench supplementary"; timeout 600 python bench.py --workload point_mul > $OUT/bench_point_mul.json 2>> $OUT/bench.err; cat $OUT/bench_point_mul.json; timeout 600 python bench.py --workload point_mul --field bn254_fr > $OUT/bench_point_mul_bn254.json 2>> $OUT/bench.err; cat $OUT/bench_point_mul_bn254.json; timeout 600 python bench.py --workload inner_product > $OUT/bench_inner_product.json 2>> $OUT/bench.err; cat $OUT/bench_inner_product.json
echo "== bench_points"; timeout 900 python tools/bench_points.py 18 > $OUT/bench_points.txt 2>&1; tail -20 $OUT/bench_points.txt
echo "== ncu launches"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
  python bench.py --steps 2 --warmup 3 --e2e-steps 0 --no-cpu-baseline > $OUT/ncu_launch_bench.log 2>&1; echo "ncu launches rc=$?"
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:beaver_recombine -s 4 -c 2 -o $OUT/prof_recombine -f \
  python bench.py --steps 2 --warmup 3 --e2e-steps 0 --no-cpu-baseline > $OUT/ncu_full_bench.log 2>&1; echo "ncu full rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:beaver_mask -s 4 -c 2 -o $OUT/prof_mask -f \
  python bench.py --steps 2 --warmup 3 --e2e-steps 0 --no-cpu-baseline > $OUT/ncu_full_mask.log 2>&1; echo "ncu full mask rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pt_beaver_recombine -c 2 -o $OUT/prof_pt_recombine -f \
  python tools/bench_points.py 17 > $OUT/ncu_full_pt.log 2>&1; echo "ncu full pt rc=$?"
ls -la $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"fr_ntt_tile|fr_ntt_strided|fr_batch_inverse" -c 4 -o $OUT/prof_ntt -f \
  python tools/bench_extra.py > $OUT/ncu_full_ntt.log 2>&1; echo "ncu full ntt rc=$?"
ls -la $OUT | head -40
