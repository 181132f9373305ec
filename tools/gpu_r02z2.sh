#!/bin/bash
# round-2 GPU visit Z2 (1 GPU): inversion with the fused top (tree up + safegcd + tree down in one kernel)
TAG=${1:-r02z2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest ntt"; timeout 900 python -m pytest tests/test_gpu_ntt.py tests/test_gpu_offline.py -x -q -m gpu 2>&1 | tail -2
echo "== bench_extra"; timeout 600 python tools/bench_extra.py > $OUT/bench_extra.txt 2>&1; grep -E "^---|inverse" $OUT/bench_extra.txt
echo "== ncu inverse"
timeout 900 ncu --set full --clock-control none -k regex:"fr_inv_" -c 6 -o /tmp/prof_inv -f python tools/bench_ntt_once.py > $OUT/ncu_full_inv.log 2>&1; echo "ncu rc=$?"
python tools/ncu_summary.py full /tmp/prof_inv.ncu-rep > $OUT/inverse_full.txt 2>&1
grep -E "^## |time_duration" $OUT/inverse_full.txt | paste - - | awk '{print $3, $(NF-1), $NF}'
echo "== memcheck inverse"; timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_ntt.py -x -q -m gpu -k "batch_inverse" > $OUT/memcheck_inv.log 2>&1; echo "memcheck rc=$?"; tail -3 $OUT/memcheck_inv.log
