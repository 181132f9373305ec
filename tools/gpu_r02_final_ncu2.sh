#!/bin/bash
# round-2: ncu evidence for the final build — point recombine (both curves), NTT and inversion kernels (summarised on the box)
TAG=${1:-r02zy}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for c in ed25519 bn254; do
  arg=""; [ $c = ed25519 ] && arg="ed25519"
  timeout 900 ncu --set full --clock-control none -k regex:"pt_beaver_recombine" -c 1 -o /tmp/prof_pt_$c -f python tools/bench_pt_bn_once.py 17 $arg > $OUT/ncu_pt_$c.log 2>&1; echo "ncu $c rc=$?"
  python tools/ncu_summary.py full /tmp/prof_pt_$c.ncu-rep > $OUT/pt_recombine_${c}_full.txt 2>&1
  ncu -i /tmp/prof_pt_$c.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
hdr=rows[0]; vals=rows[2] if len(rows)>2 else rows[1]
for h,v in zip(hdr,vals):
    if ('inst_executed_pipe_fma' in h or 'inst_executed_pipe_alu' in h or 'no_instruction' in h) and 'avg' in h: print(h, v)
" >> $OUT/pt_recombine_${c}_full.txt 2>&1
done
timeout 900 ncu --set full --clock-control none -k regex:"fr_ntt_|fr_inv_" -c 12 -o /tmp/prof_ntt -f python tools/bench_ntt_once.py > $OUT/ncu_ntt.log 2>&1; echo "ncu ntt rc=$?"
python tools/ncu_summary.py full /tmp/prof_ntt.ncu-rep > $OUT/ntt_inverse_full.txt 2>&1
grep -E "^## |time_duration" $OUT/ntt_inverse_full.txt | paste - - | awk '{print $3, $(NF-1), $NF}'
grep -E "^## |time_duration|inst_executed.sum|pipe_fma" $OUT/pt_recombine_*_full.txt | cut -c1-200
