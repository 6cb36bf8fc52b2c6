set -x
cd $GRAFT_REPO_ROOT
( time timeout 1500 python -m pytest tests/test_gpu_multi.py tests/test_multigrid.py tests/test_gpu_mfem_shim.py -m gpu -q -x 2>&1 | tail -25 ) > gpurun_out/r2s_pytest.log 2>&1
cat gpurun_out/r2s_pytest.log
