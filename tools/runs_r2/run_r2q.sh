set -x
cd $GRAFT_REPO_ROOT
nvidia-smi topo -m 2>&1 | head -30 > gpurun_out/r2q_topo.txt
for d in /sys/bus/pci/devices/*; do c=$(cat $d/class 2>/dev/null); case "$c" in 0x0302*|0x0300*) echo "$d $(cat $d/numa_node 2>/dev/null) $(cat $d/vendor)";; esac; done >> gpurun_out/r2q_topo.txt 2>&1
ls /sys/devices/system/node/ >> gpurun_out/r2q_topo.txt 2>&1; nproc >> gpurun_out/r2q_topo.txt
cat gpurun_out/r2q_topo.txt
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 --legs parity --no-cpu > gpurun_out/r2q_bench_n8.json 2> gpurun_out/r2q_bench_n8.err ) 2> gpurun_out/r2q_bench_n8.time
cat gpurun_out/r2q_bench_n8.time; grep -v '^$' gpurun_out/r2q_bench_n8.err | grep -v 'OMP_NUM\|^\*\*\*' | tail -5
python -c "
import json; d=json.loads(open('gpurun_out/r2q_bench_n8.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e'], d.get('e2e_pcg'))"
