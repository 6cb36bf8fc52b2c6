#!/bin/bash
# parity + bench (apply only) for every tuning variant of one order: tools/tune_run.sh P N
P=$1; N=$2
D=$((P+1))
for so in cardiac-ablation-ecm2_b200/libb200pa_d${D}_*.so; do
  ok=$(B200PA_LIB=$PWD/$so timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bioheat.py -m gpu -q -x -k "(form_mult or pcg or full_size) and (p$P or $P-)" 2>&1 | tail -1)
  B200PA_LIB=$PWD/$so timeout 300 python bench.py --order $P --n $N --steps 20 --warmup 3 --no-cpu --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$so'.split('libb200pa_')[1], round(d['value'],2),'GDOF/s elem_ms',round(d['roofline']['ms_per_launch'],4),'elem_frac',round(d['roofline']['frac'],3),'apply_frac',round(d['roofline_apply']['frac'],3), '| parity: $ok')"
done
