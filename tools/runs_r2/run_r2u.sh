set -x
cd $GRAFT_REPO_ROOT
bash tools/tune_run.sh 2>&1 | grep -v "^+" > gpurun_out/r2u_tuning_two_stage_stored.txt
cat gpurun_out/r2u_tuning_two_stage_stored.txt
( time timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -12 ) > gpurun_out/r2u_pytest.log 2>&1
cat gpurun_out/r2u_pytest.log
