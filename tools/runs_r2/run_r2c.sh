set -x
cd $GRAFT_REPO_ROOT
ncu --set full --clock-control none --import-source on -k regex:"k_diag_sf|pa_element_kernel" -c 10 -f -o gpurun_out/r2c_prof_setup python tools/prof_setup.py > gpurun_out/r2c_ncu_setup.log 2>&1
tail -3 gpurun_out/r2c_ncu_setup.log
for pn in "4 50" "5 40" "6 34"; do set -- $pn
  ncu --set full --clock-control none --import-source on -k regex:"pa_apply_kernel" -c 1 -f -o gpurun_out/r2c_prof_apply_p$1 python bench.py --order $1 --elems $2 --steps 1 --warmup 3 --no-cpu --no-extras > gpurun_out/r2c_ncu_apply_p$1.log 2>&1
  tail -2 gpurun_out/r2c_ncu_apply_p$1.log
done
ls -la gpurun_out/*.ncu-rep
