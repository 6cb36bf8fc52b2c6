set -x
cd $GRAFT_REPO_ROOT
( time timeout 600 python bench.py --gpus 1 > gpurun_out/r2w_bench_n1.json 2> gpurun_out/r2w_bench_n1.err ) 2> gpurun_out/r2w_bench_n1.time
cat gpurun_out/r2w_bench_n1.time; tail -c 300 gpurun_out/r2w_bench_n1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2w_launches_bench.csv python bench.py --steps 5 --warmup 3 --no-cpu --no-extras > gpurun_out/r2w_ncu1.log 2>&1; tail -2 gpurun_out/r2w_ncu1.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"pa_apply_kernel|k_segment_sum" --launch-skip 6 -c 4 -f -o gpurun_out/r2w_prof_apply_p2 python bench.py --steps 2 --warmup 3 --no-cpu --no-extras > gpurun_out/r2w_ncu2.log 2>&1; tail -2 gpurun_out/r2w_ncu2.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_diag_sf" --launch-skip 4 -c 3 -f -o gpurun_out/r2w_prof_diag python tools/setup_bench.py > gpurun_out/r2w_ncu3.log 2>&1; tail -2 gpurun_out/r2w_ncu3.log
ls -la gpurun_out/r2w_*
