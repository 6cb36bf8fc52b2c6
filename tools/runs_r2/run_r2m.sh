set -x
cd $GRAFT_REPO_ROOT
( time timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -12 ) > gpurun_out/r2m_pytest.log 2>&1
cat gpurun_out/r2m_pytest.log
( time timeout 900 python bench.py > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err ) 2> gpurun_out/r2m_bench.time
cat gpurun_out/r2m_bench.time; tail -c 300 gpurun_out/r2m_bench.err
