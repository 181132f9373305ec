#!/bin/bash
# round-2 GPU visit Z14 (1 GPU): L2 eviction priorities in K1 / K2 (read-once operands evict-first, masks evict-last)
OUT=gpurun_out/r02z14; mkdir -p $OUT
for h in 0 1; do for st in 200 20; do
  echo "== ARKMPC_L2_HINT=$h steps $st"; ARKMPC_L2_HINT=$h timeout 300 python bench.py --steps $st --warmup 5 --configs none --e2e-steps 0 --no-cpu-baseline > $OUT/bench_h${h}_$st.json 2>> $OUT/bench.err
  python -c "import json;d=json.load(open('$OUT/bench_h${h}_$st.json'));print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_us'], d['clocks']['sm_mhz'], d['clocks']['reasons'])"
done; done
echo "== 2^22 and c25519 with hints"; for h in 0 1; do ARKMPC_L2_HINT=$h timeout 300 python bench.py --steps 100 --log2-batch 22 --configs none --e2e-steps 0 --no-cpu-baseline 2>> $OUT/bench.err | python -c "import json,sys;d=json.loads(sys.stdin.read());print('2^22 hint $h', d['value'], d['roofline']['frac'])"; done
echo "== parity with hints"; ARKMPC_L2_HINT=1 timeout 900 python -m pytest tests/test_gpu_fr.py -x -q -m gpu -k "beaver or hint or host_buffer" 2>&1 | tail -2
echo "== ncu K2 with hints"
ARKMPC_L2_HINT=1 timeout 900 ncu --set full --clock-control none -k regex:"beaver_recombine_kernel|beaver_mask" -s 8 -c 4 -o /tmp/prof_l2 -f python bench.py --steps 2 --warmup 3 --e2e-steps 0 --no-cpu-baseline --configs none > $OUT/ncu.log 2>&1; echo "ncu rc=$?"
python tools/ncu_summary.py full /tmp/prof_l2.ncu-rep > $OUT/l2hint_full.txt 2>&1; grep -E "^## |time_duration|dram__bytes" $OUT/l2hint_full.txt | head -16
