set -x
cd $GRAFT_REPO_ROOT
( time timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 ) > gpurun_out/r2zz_pytest_all_2gpu.log 2>&1
cat gpurun_out/r2zz_pytest_all_2gpu.log
: > gpurun_out/r2zz_sweep_big_sizes.jsonl
for pn in "2 200" "3 134" "4 100" "2 290"; do
  set -- $pn
  CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --order $1 --elems $2 --ops both --steps 10 --warmup 3 --no-cpu --no-extras 2>/dev/null >> gpurun_out/r2zz_sweep_big_sizes.jsonl
done
python - <<'P'
import json
for l in open("gpurun_out/r2zz_sweep_big_sizes.jsonl"):
    d = json.loads(l); c = d["config"]
    print(c["order"], c["dofs_per_gpu"], round(d["value"], 2), "GDOF/s apply_frac", round(d["roofline_apply"]["frac"], 3), "kernel", round(d["roofline"]["frac"], 3), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
P
