set -x
cd $GRAFT_REPO_ROOT
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bioheat.py tests/test_gpu_factorised.py -m gpu -q -x 2>&1 | tail -8 ) > gpurun_out/r2l_pytest.log 2>&1
cat gpurun_out/r2l_pytest.log
for dm in 1 0; do for ops in both diff; do
  B200PA_DMMA=$dm timeout 300 python bench.py --order 6 --elems 34 --ops $ops --steps 20 --warmup 3 --no-cpu --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('dmma=$dm ops=$ops', round(d['value'],2),'GDOF/s kern_ms',round(d['roofline']['ms_per_launch'],4),'kern_frac',round(d['roofline']['frac'],3),'apply_frac',round(d['roofline_apply']['frac'],3))" | tee -a gpurun_out/r2l_dmma_p6.txt
done; done
ncu --set full --clock-control none --import-source on -k regex:"pa_apply_dmma" -c 1 -f -o gpurun_out/r2l_prof_apply_p6_dmma python bench.py --order 6 --elems 34 --steps 1 --warmup 3 --no-cpu --no-extras > gpurun_out/r2l_ncu.log 2>&1
tail -2 gpurun_out/r2l_ncu.log
