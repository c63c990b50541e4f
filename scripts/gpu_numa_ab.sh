for f in "" "--no-numa-bind" ""; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu --no-others --no-sustained $f 2>/dev/null | grep "^{" | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('[$f]', d['n_gpus'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e'].get('host_numa_node'))"
done
python bench.py --no-cpu --no-others --no-sustained 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('1 gpu e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e'].get('host_numa_node'))"
nvidia-smi topo -m | head -12; lscpu | grep -i "numa\|socket\|^CPU(s)"
