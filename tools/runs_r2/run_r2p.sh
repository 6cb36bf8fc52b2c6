set -x
cd $GRAFT_REPO_ROOT
( time timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -12 ) > gpurun_out/r2p_pytest.log 2>&1
cat gpurun_out/r2p_pytest.log
timeout 300 python tools/step_launches.py > gpurun_out/r2p_step.json 2> gpurun_out/r2p_step.err; tail -3 gpurun_out/r2p_step.err; cat gpurun_out/r2p_step.json
( time timeout 900 python bench.py > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err ) 2> gpurun_out/r2p_bench.time
cat gpurun_out/r2p_bench.time; tail -c 300 gpurun_out/r2p_bench.err
