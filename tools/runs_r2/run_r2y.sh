set -x
cd $GRAFT_REPO_ROOT
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r2y_bench_n4.json 2> gpurun_out/r2y_bench_n4.err ) 2> gpurun_out/r2y_bench_n4.time
cat gpurun_out/r2y_bench_n4.time; grep -v '^$' gpurun_out/r2y_bench_n4.err | grep -v 'OMP_NUM\|^\*\*\*' | tail -5
