# usage: bash scripts/gpu_c5_sweep.sh N "<patches per GPU ...>"  -- BASELINE config C5: 3-D Euler batch-size sweep on N GPUs
# (N > 1: run under gpurun --gpus N; the all-reduce(max) of lambda runs in the patch kernel's epilogue)
N=$1
for B in $2; do
  if [ "$N" = 1 ]; then
    python bench.py --batch $B --steps 50 --no-cpu --no-e2e --no-others --no-sustained 2>/dev/null | tail -1 > gpurun_out/c5_${N}_$B.json
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
      bench.py --gpus $N --batch $B --steps 50 --warmup 5 --no-cpu --no-e2e 2>/dev/null | grep '^{' | tail -1 > gpurun_out/c5_${N}_$B.json
  fi
  python - <<PY
import json
d=json.loads(open("gpurun_out/c5_${N}_$B.json").read())
print(f"gpus {d['n_gpus']} patches/gpu {d['config']['patches_per_gpu']:7d} ms/step {d['ms_per_step']:.4f} cell-updates/s {d['value']:.3e} "
      f"alg GB/s per GPU {d['roofline']['achieved']:.0f} frac {d['roofline']['frac']:.3f}")
PY
done
