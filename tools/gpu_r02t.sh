#!/bin/bash
# round-2 GPU visit T (1 GPU): BN254 G1 point recombine — ncu capture of the shipped kernel and the 384-thread variant
# (the reports are large: they are summarised on the box and not brought back)
TAG=${1:-r02t}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for b in 256 384; do
  echo "== ncu bn254 recombine ($b)"
  ARKMPC_PT_BN_BLOCK=$b timeout 900 ncu --set full --clock-control none -k regex:"pt_beaver_recombine" -c 1 -o /tmp/prof_bn$b -f python tools/bench_pt_bn_once.py 17 > $OUT/ncu_bn$b.log 2>&1; echo "ncu rc=$?"
  python tools/ncu_summary.py full /tmp/prof_bn$b.ncu-rep > $OUT/bn254_pt_recombine_$b.txt 2>&1
  ncu -i /tmp/prof_bn$b.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
hdr=rows[0]; vals=rows[2] if len(rows)>2 else rows[1]
for h,v in zip(hdr,vals):
    if 'stalled' in h and 'per_issue_active' not in h and 'pct' in h or 'inst_executed_pipe' in h or 'no_instruction' in h or 'imc' in h.lower() or 'icc' in h.lower():
        print(h, v)
" > $OUT/bn254_pt_recombine_${b}_stalls.txt 2>&1
done
ls -la $OUT
