set -x
cd $GRAFT_REPO_ROOT
./oracle/_ref/shim_check surface 2 4 > gpurun_out/r2e_surface.log 2>&1; echo "rc=$?" >> gpurun_out/r2e_surface.log
./oracle/_ref/shim_check rf 2 4 2 > gpurun_out/r2e_rf.log 2>&1; echo "rc=$?" >> gpurun_out/r2e_rf.log
tail -30 gpurun_out/r2e_surface.log gpurun_out/r2e_rf.log
( time timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 ) > gpurun_out/r2e_pytest.log 2>&1
cat gpurun_out/r2e_pytest.log
for so in libb200pa.so libb200pa_diag44.so libb200pa_diag90.so; do
  echo "== $so" >> gpurun_out/r2e_setup.jsonl
  B200PA_LIB=$PWD/cardiac-ablation-ecm2_b200/$so timeout 300 python tools/setup_bench.py >> gpurun_out/r2e_setup.jsonl 2>> gpurun_out/r2e_setup.err
done
ncu --set full --clock-control none --import-source on -k regex:"k_diag_sf" -c 2 -f -o gpurun_out/r2e_prof_diag python tools/prof_setup.py > gpurun_out/r2e_ncu.log 2>&1
