#!/bin/bash
# quick order sweep (diffusion+mass apply only): orders 1..6 at ~8M dofs + two larger points.
# usage (GPU box): bash tools/sweep_quick.sh > gpurun_out/sweep.jsonl
for pn in "1 200" "2 100" "3 67" "4 50" "5 40" "6 34" "2 200" "3 134"; do
  set -- $pn
  timeout 600 python bench.py --order $1 --elems $2 --ops both --steps 20 --warmup 3 --no-cpu --no-extras 2>/dev/null
done
