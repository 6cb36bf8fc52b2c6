set -x
cd $GRAFT_REPO_ROOT
( timeout 900 python -m pytest tests/test_gpu_lifecycle.py -m gpu -q -x 2>&1 | tail -15 ) > gpurun_out/r2t_pytest.log 2>&1
cat gpurun_out/r2t_pytest.log
bash tools/ref_cuda_bench.sh r2t 2>&1 | grep -v "^+" | cut -c1-600
tail -5 gpurun_out/r2t_ref_cuda.err
