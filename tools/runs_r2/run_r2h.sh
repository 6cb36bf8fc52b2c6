set -x
cd $GRAFT_REPO_ROOT
( time timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 ) > gpurun_out/r2h_pytest.log 2>&1
cat gpurun_out/r2h_pytest.log
for so in libb200pa.so $(cd cardiac-ablation-ecm2_b200; ls libb200pa_diag_*.so); do
  echo "== $so" >> gpurun_out/r2h_setup.jsonl
  B200PA_LIB=$PWD/cardiac-ablation-ecm2_b200/$so timeout 300 python tools/setup_bench.py >> gpurun_out/r2h_setup.jsonl 2>> gpurun_out/r2h_setup.err
done
for so in libb200pa.so libb200pa_diag_g4k44.so libb200pa_diag_g7k44.so; do
  echo "== $so p3" >> gpurun_out/r2h_setup.jsonl
  B200PA_LIB=$PWD/cardiac-ablation-ecm2_b200/$so timeout 300 python tools/setup_bench.py --order 3 --elems 67 >> gpurun_out/r2h_setup.jsonl 2>> gpurun_out/r2h_setup.err
done
timeout 600 python bench.py --legs bioheat --no-cpu > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err
tail -c 300 gpurun_out/r2h_setup.err gpurun_out/r2h_bench.err
