#!/bin/bash
# bench (apply only, ~8M dofs) every tuning variant libb200pa_d<D>_<tag>.so next to the default build;
# usage (GPU box): [QDATA=factorised] bash tools/tune_run.sh [N_for_p2]   (parity of the chosen variant is checked by the test-suite afterwards)
declare -A NEL=([2]=200 [3]=100 [4]=67 [5]=50 [6]=40 [7]=34)
[ -n "$1" ] && NEL[3]=$1
for so in cardiac-ablation-ecm2_b200/libb200pa.so $(ls cardiac-ablation-ecm2_b200/libb200pa_d*.so 2>/dev/null); do
  name=$(basename $so .so)
  if [ "$name" == "libb200pa" ]; then ds="2 3 4 5 6 7"; else ds=$(echo $name | sed 's/libb200pa_d\([0-9]\)_.*/\1/'); fi
  for D in $ds; do
    P=$((D-1))
    B200PA_LIB=$PWD/$so timeout 300 python bench.py --order $P --elems ${NEL[$D]} --steps 20 --warmup 3 --no-cpu --no-extras --qdata ${QDATA:-stored} 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$name p=$P', round(d['value'],2),'GDOF/s kern_ms',round(d['roofline']['ms_per_launch'],4),'kern_frac',round(d['roofline']['frac'],3),'apply_frac',round(d['roofline_apply']['frac'],3))"
  done
done
