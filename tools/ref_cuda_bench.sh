#!/bin/bash
# The reference's OWN CUDA backend (oracle/_ref/ref_driver_cuda = the unmodified reference sources compiled with nvcc -x cu
# for sm_100, `make -C oracle refcuda`) on the same B200: PA diffusion+mass apply L->L and a 20-iteration Jacobi-PCG through
# BilinearForm / CGSolver with Device("cuda"), next to this library on the same workloads.  Test infrastructure (a second
# baseline); writes gpurun_out/<tag>_ref_cuda.jsonl.
#   bash tools/ref_cuda_bench.sh <tag>
set -u
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
tag=${1:-refcuda}
out=gpurun_out/${tag}_ref_cuda.jsonl
: > $out
for pn in "2 100" "1 128" "3 67" "4 50" "5 40" "6 34"; do
  set -- $pn
  timeout 900 oracle/_ref/ref_driver_cuda time_apply $1 $2 20 3 cuda 20 >> $out 2>> gpurun_out/${tag}_ref_cuda.err || echo "{\"failed\": \"p=$1 N=$2\"}" >> $out
  tail -1 $out | cut -c1-400
  timeout 600 python bench.py --order $1 --elems $2 --steps 20 --warmup 3 --no-cpu --legs bioheat > gpurun_out/${tag}_b200_p$1.json 2>> gpurun_out/${tag}_ref_cuda.err
  python - <<P
import json
d = json.loads(open("gpurun_out/${tag}_b200_p$1.json").read().strip().splitlines()[-1])
print("b200pa p=$1 N=$2:", round(d["ms_per_step"], 4), "ms/apply", round(d["value"], 2), "GDOF/s; pcg ms/it", d.get("pcg", {}).get("ms_per_iter"))
P
done
