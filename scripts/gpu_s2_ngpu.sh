# usage: bash scripts/gpu_s2_ngpu.sh N -- the driver's command line on N GPUs with the final build + the C5 small batch
N=$1
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/s2_${N}gpu_default.json 2> gpurun_out/s2_${N}gpu_default.err; echo default rc=$?
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $N --batch 4096 --steps 50 --warmup 5 --no-cpu --no-e2e > gpurun_out/s2_${N}gpu_b4096.json 2> gpurun_out/s2_${N}gpu_b4096.err; echo b4096 rc=$?
python - <<PY
import json
for name in ("default","b4096"):
    d=[json.loads(l) for l in open("gpurun_out/s2_${N}gpu_%s.json" % name) if l.startswith("{")][-1]
    print(name, d["n_gpus"], "ms/step %.4f" % d["ms_per_step"], "value %.4e" % d["value"], "frac/GPU %.3f" % d["roofline"]["frac"], "bitwise", d.get("multi_gpu_bitwise"), "e2e ms", (d.get("e2e") or {}).get("ms_per_step"))
PY
