# usage: bash scripts/gpu_multi.sh N  -- N-GPU bench: all-reduce fused into the patch kernel, as a separate peer-memory
# kernel, and through NCCL (run under gpurun --gpus N)
N=$1
for red in "peer" "peer --no-fused" "nccl"; do
  tag=$(echo $red | tr -d ' -')
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 50 --warmup 5 --no-cpu --no-e2e --reducer $red > gpurun_out/multi_${N}_$tag.json 2> gpurun_out/multi_${N}_$tag.err
  echo "$red rc=$?"; tail -2 gpurun_out/multi_${N}_$tag.err
  python - <<PY
import json
try:
    d=[json.loads(l) for l in open("gpurun_out/multi_${N}_$tag.json") if l.startswith("{")][0]
    print("$red", d["n_gpus"], "ms/step", round(d["ms_per_step"],4), "kernel_ms", round(d["roofline"]["kernel_ms"],4), d["gpu_launches"], d["config"]["parallelism"])
except Exception as e: print("no result", e)
PY
done
