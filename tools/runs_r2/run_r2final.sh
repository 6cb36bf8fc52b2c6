set -x
cd $GRAFT_REPO_ROOT
( time timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 ) > gpurun_out/r2final_pytest.log 2>&1
cat gpurun_out/r2final_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time timeout 600 python bench.py > gpurun_out/r2final_bench_n1.json 2> gpurun_out/r2final_bench_n1.err ) 2> gpurun_out/r2final_bench_n1.time
cat gpurun_out/r2final_bench_n1.time; tail -c 300 gpurun_out/r2final_bench_n1.err; cut -c1-400 gpurun_out/r2final_bench_n1.json
