#!/bin/bash
# configs[3] upper end: large single-GPU problems (J-free geometry keeps 200 M dofs at p=2 within 180 GB)
for pn in "2 200" "2 290" "3 134" "3 190" "1 400" "4 100"; do
  set -- $pn
  timeout 900 python bench.py --order $1 --elems $2 --ops both --steps 10 --warmup 3 --no-cpu --no-extras 2>/dev/null
done
