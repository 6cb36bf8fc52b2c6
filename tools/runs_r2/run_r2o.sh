set -x
cd $GRAFT_REPO_ROOT
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "host" 2>&1 | tail -5 ) > gpurun_out/r2o_pytest.log 2>&1
cat gpurun_out/r2o_pytest.log
timeout 300 python tools/step_launches.py > gpurun_out/r2o_step.json 2> gpurun_out/r2o_step.err; tail -3 gpurun_out/r2o_step.err; cat gpurun_out/r2o_step.json
timeout 300 python tools/step_launches.py --factorised > gpurun_out/r2o_step_fact.json 2>> gpurun_out/r2o_step.err; cat gpurun_out/r2o_step_fact.json
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2o_step_launches.csv python tools/step_launches.py > gpurun_out/r2o_step_ncu.log 2>&1
tail -2 gpurun_out/r2o_step_ncu.log
python - <<'P'
import csv, collections
rows = [r for r in csv.reader(l for l in open("gpurun_out/r2o_step_launches.csv") if not l.startswith("=="))]
h = rows[0]; ki, vi = h.index("Kernel Name"), h.index("Metric Value")
seq = [(r[ki].split("(")[0][-60:], float(r[vi].replace(",", ""))) for r in rows[1:] if len(r) > vi and r[vi]]
tot = collections.OrderedDict()
for k, v in seq:
    tot.setdefault(k, [0, 0.0]); tot[k][0] += 1; tot[k][1] += v
print("launches", len(seq), "sum_us", sum(v for _, v in seq) / 1e3)
for k, (n, v) in tot.items():
    print(f"{n:4d} {v/1e3:10.1f} us  {k}")
P
