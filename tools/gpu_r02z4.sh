#!/bin/bash
# round-2 GPU visit Z4 (1 GPU): cooperative sweeps for the small levels of the inversion tree
TAG=${1:-r02z4}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest ntt (default)"; timeout 900 python -m pytest tests/test_gpu_ntt.py tests/test_gpu_offline.py -x -q -m gpu 2>&1 | tail -2
echo "== pytest ntt (no single-launch path: cooperative sweeps at every small size)"; ARKMPC_INV_SMALL=0 timeout 900 python -m pytest tests/test_gpu_ntt.py -x -q -m gpu -k inverse 2>&1 | tail -2
for cfg in "ARKMPC_INV_COOP=1 ARKMPC_INV_SMALL=1" "ARKMPC_INV_COOP=0 ARKMPC_INV_SMALL=1" "ARKMPC_INV_COOP=1 ARKMPC_INV_SMALL=0"; do
  echo "== $cfg"; env $cfg timeout 600 python tools/bench_extra.py 2>&1 | grep -E "^---|inverse"
done
echo "== ncu inverse"
timeout 900 ncu --set full --clock-control none -k regex:"fr_inv_" -c 6 -o /tmp/prof_inv -f python tools/bench_ntt_once.py > $OUT/ncu_full_inv.log 2>&1; echo "ncu rc=$?"
python tools/ncu_summary.py full /tmp/prof_inv.ncu-rep > $OUT/inverse_full.txt 2>&1
grep -E "^## |time_duration" $OUT/inverse_full.txt | paste - - | awk '{print $3, $(NF-1), $NF}'
echo "== memcheck inverse"; timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_ntt.py -x -q -m gpu -k "batch_inverse" > $OUT/memcheck_inv.log 2>&1; echo "memcheck rc=$?"; tail -3 $OUT/memcheck_inv.log
echo "== racecheck inverse"; timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_ntt.py -x -q -m gpu -k "batch_inverse and 131073" > $OUT/racecheck_inv.log 2>&1; echo "racecheck rc=$?"; tail -3 $OUT/racecheck_inv.log
