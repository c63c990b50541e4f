# usage: bash scripts/gpu_r02_multi_n.sh N  -- everything the round needs from an N-GPU box in one call
set -x
N=$1
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02_topo_${N}gpu.txt 2>&1
run() {  # name, extra bench args
  name=$1; shift
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 50 --warmup 5 --no-cpu --no-e2e --trace gpurun_out/r02_trace_${N}_$name "$@" \
    > gpurun_out/r02_multi_${N}_$name.json 2> gpurun_out/r02_multi_${N}_$name.err
  echo "$name rc=$?"; tail -2 gpurun_out/r02_multi_${N}_$name.err
  python scripts/exchange_attribution.py gpurun_out/r02_trace_${N}_$name > gpurun_out/r02_attribution_${N}_$name.txt 2>&1
}
# the same box's first GPU alone, for the weak-scaling reference
python bench.py --steps 50 --warmup 5 --no-cpu --no-e2e --no-others --no-sustained > gpurun_out/r02_multi_${N}_single.json 2>/dev/null
python -c "
import json; d=[json.loads(l) for l in open('gpurun_out/r02_multi_${N}_single.json') if l.startswith('{')][0]
print('single GPU on this box: ms/step', round(d['ms_per_step'],4), 'value %.4e' % d['value'])"
run loop --reducer peer
run blocking --reducer peer --time-step host
run nccl --reducer nccl
bash scripts/gpu_c5_sweep.sh $N "4096 16384 65536" | tee gpurun_out/r02_c5_sweep_${N}gpu.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 scripts/pcie_probe.py > gpurun_out/r02_pcie_probe_${N}gpu.txt 2>&1; echo pcie rc=$?
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err; echo bench rc=$?
python - <<PY
import json
for name in ("loop","blocking","nccl"):
    try:
        d=[json.loads(l) for l in open("gpurun_out/r02_multi_${N}_%s.json" % name) if l.startswith("{")][0]
        print(name, d["n_gpus"], "ms/step", round(d["ms_per_step"],4), "value %.4e" % d["value"], "bitwise", d.get("multi_gpu_bitwise"))
    except Exception as e: print(name, "no result", e)
try:
    d=[json.loads(l) for l in open("gpurun_out/r02_bench_${N}gpu.json") if l.startswith("{")][0]
    print("default bench", d["n_gpus"], "ms/step", round(d["ms_per_step"],4), "value %.4e" % d["value"], "e2e", d["e2e"]["ms_per_step"], "%.3e" % d["e2e"]["value"], "bitwise", d.get("multi_gpu_bitwise"))
except Exception as e: print("default no result", e)
PY
cat gpurun_out/r02_pcie_probe_${N}gpu.txt | grep -v "^\*\|OMP_NUM"
