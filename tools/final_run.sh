#!/bin/bash
# round-end measurement set on one B200: tests, bench (with the CPU reference leg), order sweeps, ncu launch list and
# one ncu --set full capture of the top kernels.  usage (GPU box): bash tools/final_run.sh <tag>
T=${1:-r1m}
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_${T}_reference.json 2>> gpurun_out/bench_$T.err
bash tools/sweep_quick.sh > gpurun_out/sweep_$T.jsonl
for pn in "1 200" "2 100" "3 67" "4 50" "5 40" "6 34"; do set -- $pn
  python bench.py --order $1 --elems $2 --qdata factorised --steps 20 --warmup 3 --no-cpu --no-extras 2>/dev/null; done > gpurun_out/sweep_${T}_factorised.jsonl
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$T.csv \
   python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/ncu_launches_$T.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"pa_apply_kernel|k_segment_sum" -c 2 -f -o gpurun_out/prof_apply_$T \
   python bench.py --steps 2 --warmup 1 --no-cpu --no-extras > gpurun_out/ncu_full_$T.log 2>&1
tail -c 600 gpurun_out/bench_$T.json
