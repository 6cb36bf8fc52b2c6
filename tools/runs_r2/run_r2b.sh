set -x
cd $GRAFT_REPO_ROOT
( time timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/r2b_pytest.log 2>&1
cat gpurun_out/r2b_pytest.log
for so in libb200pa.so libb200pa_diag40.so libb200pa_diag110.so; do
  echo "== $so" >> gpurun_out/r2b_setup.jsonl
  B200PA_LIB=$PWD/cardiac-ablation-ecm2_b200/$so timeout 300 python tools/setup_bench.py >> gpurun_out/r2b_setup.jsonl 2>> gpurun_out/r2b_setup.err
done
for pn in "1 200" "3 67" "4 50" "6 34"; do set -- $pn
  timeout 300 python tools/setup_bench.py --order $1 --elems $2 >> gpurun_out/r2b_setup.jsonl 2>> gpurun_out/r2b_setup.err
done
timeout 300 python tools/setup_bench.py --skew >> gpurun_out/r2b_setup.jsonl 2>> gpurun_out/r2b_setup.err
timeout 600 python bench.py --legs bioheat,rf,factorised --no-cpu > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
tail -c 600 gpurun_out/r2b_setup.err gpurun_out/r2b_bench.err
