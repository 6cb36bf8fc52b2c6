set -x
cd $GRAFT_REPO_ROOT
( time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > gpurun_out/r2d_pytest.log 2>&1
cat gpurun_out/r2d_pytest.log
for a in "" "--order 1 --elems 200" "--order 3 --elems 67" "--order 4 --elems 50"; do
  timeout 300 python tools/setup_bench.py $a >> gpurun_out/r2d_setup.jsonl 2>> gpurun_out/r2d_setup.err
done
timeout 600 python bench.py --legs bioheat,rf,factorised --no-cpu > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err
ncu --set full --clock-control none --import-source on -k regex:"k_diag_sf|pa_element_kernel" -c 10 -f -o gpurun_out/r2d_prof_setup python tools/prof_setup.py > gpurun_out/r2d_ncu_setup.log 2>&1
tail -c 600 gpurun_out/r2d_setup.err gpurun_out/r2d_bench.err
