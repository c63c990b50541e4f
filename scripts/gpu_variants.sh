# usage: bash scripts/gpu_variants.sh "<variant names or 'main'>" "<workloads>"  -- kernel-only timings of tuning builds
for v in $1; do
  if [ "$v" = main ]; then unset EXAHYPE_CUDA_LIB; else export EXAHYPE_CUDA_LIB=$PWD/exahype_b200/variants/$v/libexahype_cuda.so; fi
  for w in $2; do
    python bench.py --workload $w --no-cpu --no-e2e --no-others --no-sustained --steps 30 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$v', '$w', round(d['roofline']['kernel_ms'],4), round(d['roofline']['achieved']), round(d['roofline']['frac'],3), d['config']['kernel'])"
  done
done
