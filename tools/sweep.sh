#!/bin/bash
# configs[3]: PA operator-apply roofline sweep, orders 1..6, ~8M dofs (and larger points) on one B200.
# usage (GPU box): bash tools/sweep.sh > gpurun_out/sweep.jsonl
for ops in both diff; do
  for pn in "1 200" "2 100" "3 67" "4 50" "5 40" "6 34" "2 200" "3 134" "1 100" "2 50"; do
    set -- $pn
    timeout 600 python bench.py --order $1 --elems $2 --ops $ops --steps 20 --warmup 3 --no-cpu --no-extras 2>/dev/null
  done
done
