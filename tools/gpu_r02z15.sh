#!/bin/bash
# round-2 GPU visit Z15 (1 GPU): L2 eviction priorities compiled into K1 / K2 (no run-time switch)
OUT=gpurun_out/r02z15; mkdir -p $OUT
for st in 200 20 20; do
  timeout 300 python bench.py --steps $st --warmup 5 --configs none --e2e-steps 0 --no-cpu-baseline > $OUT/bench_$st.json 2>> $OUT/bench.err
  python -c "import json;d=json.load(open('$OUT/bench_$st.json'));print($st, d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_us'], d['clocks']['sm_mhz'], d['clocks']['reasons'])"
done
timeout 300 python bench.py --steps 100 --log2-batch 22 --configs none --e2e-steps 0 --no-cpu-baseline 2>> $OUT/bench.err | python -c "import json,sys;d=json.loads(sys.stdin.read());print('2^22', d['value'], d['roofline']['frac'])"
echo "== parity"; timeout 900 python -m pytest tests/test_gpu_fr.py -x -q -m gpu -k "beaver or hint or host_buffer" 2>&1 | tail -2
