set -x
cd $GRAFT_REPO_ROOT
( timeout 900 python -m pytest tests/test_gpu_mfem_shim.py -m gpu -q -x -k "example or surface" 2>&1 | tail -15 ) > gpurun_out/r2zy_pytest.log 2>&1; cat gpurun_out/r2zy_pytest.log
mkdir -p /tmp/ex && cd /tmp/ex && timeout 300 $GRAFT_REPO_ROOT/oracle/_ref/rf_ablation -n 32 -o 2 -dt 0.5 -tf 5 -vs 2 > $GRAFT_REPO_ROOT/gpurun_out/r2zy_example_n32.txt 2>&1; tail -30 $GRAFT_REPO_ROOT/gpurun_out/r2zy_example_n32.txt | cut -c1-200
