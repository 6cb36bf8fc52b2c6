set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L
( time timeout 900 python -m pytest tests/test_gpu_multi.py "tests/test_gpu_parity.py::test_mult_host_pipelined_equals_device_apply" -m gpu -q 2>&1 | tail -15 ) > gpurun_out/r2g_pytest.log 2>&1
cat gpurun_out/r2g_pytest.log
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2g_bench_n2.json 2> gpurun_out/r2g_bench_n2.err ) 2> gpurun_out/r2g_bench_n2.time
cat gpurun_out/r2g_bench_n2.time; tail -c 1500 gpurun_out/r2g_bench_n2.err
( time timeout 600 python bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/r2g_ref_n2.json 2>&1 ) 2>&1 | tail -3
free -g | head -2
