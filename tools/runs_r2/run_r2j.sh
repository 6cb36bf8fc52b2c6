set -x
cd $GRAFT_REPO_ROOT
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2j_bench_n2.json 2> gpurun_out/r2j_bench_n2.err ) 2> gpurun_out/r2j_bench_n2.time
cat gpurun_out/r2j_bench_n2.time; grep -v "^$" gpurun_out/r2j_bench_n2.err | grep -v "OMP_NUM\|^\*\*\*" | tail -20
CUDA_VISIBLE_DEVICES=0 timeout 300 python tools/e2e_sweep.py > gpurun_out/r2j_e2e_sweep.json 2> gpurun_out/r2j_e2e_sweep.err
cat gpurun_out/r2j_e2e_sweep.json
