set -x
cd $GRAFT_REPO_ROOT
( time timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 ) > gpurun_out/r2f_pytest.log 2>&1
cat gpurun_out/r2f_pytest.log
timeout 300 python tools/setup_bench.py >> gpurun_out/r2f_setup.jsonl 2>> gpurun_out/r2f_setup.err
timeout 300 python tools/setup_bench.py --order 3 --elems 67 >> gpurun_out/r2f_setup.jsonl 2>> gpurun_out/r2f_setup.err
( time timeout 900 python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err ) 2> gpurun_out/r2f_bench.time
cat gpurun_out/r2f_bench.time; tail -c 400 gpurun_out/r2f_bench.err
ncu --set full --clock-control none --import-source on -k regex:"k_diag_sf" -c 2 -f -o gpurun_out/r2f_prof_diag python tools/prof_setup.py > gpurun_out/r2f_ncu.log 2>&1
