set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,memory.total --format=csv
( time timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/r2a_pytest.log 2>&1
( time timeout 900 python bench.py > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err )  2> gpurun_out/r2a_bench.time
tail -c 1500 gpurun_out/r2a_bench.err
cat gpurun_out/r2a_pytest.log gpurun_out/r2a_bench.time
