set -x
cd $GRAFT_REPO_ROOT
( time timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 ) > gpurun_out/r2i_pytest.log 2>&1
cat gpurun_out/r2i_pytest.log
timeout 300 python tools/setup_bench.py > gpurun_out/r2i_setup.jsonl 2> gpurun_out/r2i_setup.err
timeout 300 python tools/e2e_sweep.py > gpurun_out/r2i_e2e_sweep.json 2> gpurun_out/r2i_e2e_sweep.err
cat gpurun_out/r2i_e2e_sweep.json
./tools/probes/dmma_probe > gpurun_out/r2i_dmma_probe.json 2>&1
cat gpurun_out/r2i_dmma_probe.json
tail -c 300 gpurun_out/r2i_setup.err gpurun_out/r2i_e2e_sweep.err
