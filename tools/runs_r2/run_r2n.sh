set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L | wc -l
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2n_bench_n8.json 2> gpurun_out/r2n_bench_n8.err ) 2> gpurun_out/r2n_bench_n8.time
cat gpurun_out/r2n_bench_n8.time; grep -v "^$" gpurun_out/r2n_bench_n8.err | grep -v "OMP_NUM\|^\*\*\*" | tail -8
( time timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -k "8-1 or 4-2" 2>&1 | tail -6 ) > gpurun_out/r2n_pytest_multi.log 2>&1
cat gpurun_out/r2n_pytest_multi.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r2n_bench_n4.json 2> gpurun_out/r2n_bench_n4.err ) 2> gpurun_out/r2n_bench_n4.time
cat gpurun_out/r2n_bench_n4.time; grep -v "^$" gpurun_out/r2n_bench_n4.err | grep -v "OMP_NUM\|^\*\*\*" | tail -5
