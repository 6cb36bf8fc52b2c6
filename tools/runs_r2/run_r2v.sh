set -x
cd $GRAFT_REPO_ROOT
( time timeout 1200 python -m pytest tests/test_gpu_multi.py tests/test_gpu_lifecycle.py tests/test_gpu_parity.py -m gpu -q -k "multi or lifecycle or host" 2>&1 | tail -6 ) > gpurun_out/r2v_pytest.log 2>&1
cat gpurun_out/r2v_pytest.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2v_bench_n2.json 2> gpurun_out/r2v_bench_n2.err ) 2> gpurun_out/r2v_bench_n2.time
cat gpurun_out/r2v_bench_n2.time; grep -v '^$' gpurun_out/r2v_bench_n2.err | grep -v 'OMP_NUM\|^\*\*\*' | tail -5
( time timeout 600 python bench.py --gpus 1 > gpurun_out/r2v_bench_n1.json 2> gpurun_out/r2v_bench_n1.err ) 2> gpurun_out/r2v_bench_n1.time
cat gpurun_out/r2v_bench_n1.time; tail -c 300 gpurun_out/r2v_bench_n1.err
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2v_bench_ref.json 2> gpurun_out/r2v_bench_ref.err ) 2> gpurun_out/r2v_bench_ref.time
cat gpurun_out/r2v_bench_ref.time; cat gpurun_out/r2v_bench_ref.json | cut -c1-600
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
