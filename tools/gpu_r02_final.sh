#!/bin/bash
# round-2 final single-GPU visit: what the driver runs (GPU suite, smoke, bench both arms) + the ncu evidence for profiles/
TAG=${1:-r02z}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > $OUT/smi.txt 2>&1
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_gpu.log
echo "== smoke"; timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke.log
echo "== bench reference"; timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; head -c 400 $OUT/bench_ref.json; echo
echo "== bench"; timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; head -c 1200 $OUT/bench.json; echo; tail -3 $OUT/bench.err
echo "== bench 2^22"; timeout 300 python bench.py --steps 200 --log2-batch 22 --configs none --e2e-steps 0 --no-cpu-baseline > $OUT/bench_2e22.json 2>> $OUT/bench.err; head -c 300 $OUT/bench_2e22.json; echo
echo "== bench c25519"; timeout 300 python bench.py --steps 200 --field curve25519_fr --configs none --e2e-steps 0 --no-cpu-baseline > $OUT/bench_c25519.json 2>> $OUT/bench.err; head -c 300 $OUT/bench_c25519.json; echo
echo "== ncu launches"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
  python bench.py --steps 2 --warmup 3 --e2e-steps 0 --no-cpu-baseline --configs none > $OUT/ncu_launch_bench.log 2>&1; echo "ncu launches rc=$?"
echo "== ncu full K2 / K1"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:beaver_recombine_kernel -s 4 -c 2 -o $OUT/prof_recombine -f \
  python bench.py --steps 2 --warmup 3 --e2e-steps 0 --no-cpu-baseline --configs none > $OUT/ncu_full_bench.log 2>&1; echo "ncu full rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:beaver_mask -s 4 -c 2 -o $OUT/prof_mask -f \
  python bench.py --steps 2 --warmup 3 --e2e-steps 0 --no-cpu-baseline --configs none > $OUT/ncu_full_mask.log 2>&1; echo "ncu full mask rc=$?"
echo "== sanitizer (memcheck, small sizes)"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_fr.py -x -q -m gpu -k "1000 or 31 or validate or hint and not 262144 and not 1048653" > $OUT/sanitizer.log 2>&1; echo "sanitizer rc=$?"; tail -4 $OUT/sanitizer.log
