set -x
cd $GRAFT_REPO_ROOT
( time timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25 ) > gpurun_out/r2k_pytest.log 2>&1
cat gpurun_out/r2k_pytest.log
timeout 600 python bench.py --legs mg,bioheat --no-cpu > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err
tail -c 600 gpurun_out/r2k_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2k_bench.json')); print(json.dumps(d.get('p_multigrid'))); print(d['e2e']); print(d['bioheat_step'])"
