# usage: bash scripts/gpu_multi.sh N [tag]  -- N-GPU bench (run under gpurun --gpus N): device-resident time loop (split-phase
# exchange), host-dt with the blocking all-reduce fused into the patch kernel / as a separate peer-memory kernel / through
# NCCL; each with the exchange's device-side stamps summarised by scripts/exchange_attribution.py
N=$1; TAG=${2:-r02}
run() {  # name, extra bench args
  name=$1; shift
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 50 --warmup 5 --no-cpu --no-e2e --trace gpurun_out/${TAG}_trace_${N}_$name "$@" \
    > gpurun_out/${TAG}_multi_${N}_$name.json 2> gpurun_out/${TAG}_multi_${N}_$name.err
  echo "$name rc=$?"; tail -2 gpurun_out/${TAG}_multi_${N}_$name.err
  python - <<PY
import json
try:
    d=[json.loads(l) for l in open("gpurun_out/${TAG}_multi_${N}_$name.json") if l.startswith("{")][0]
    print("$name", d["n_gpus"], "ms/step", round(d["ms_per_step"],4), "kernel_ms", round(d["roofline"]["kernel_ms"],4), "launches", d["gpu_launches"],
          "bitwise", d.get("multi_gpu_bitwise"), "|", d["config"]["time_step"][:40])
except Exception as e: print("no result", e)
PY
  python scripts/exchange_attribution.py gpurun_out/${TAG}_trace_${N}_$name > gpurun_out/${TAG}_attribution_${N}_$name.txt 2>&1
  cat gpurun_out/${TAG}_attribution_${N}_$name.txt | head -24
}
run loop --reducer peer
run blocking --reducer peer --time-step host
run separate --reducer peer --time-step host --no-fused
run nccl --reducer nccl
