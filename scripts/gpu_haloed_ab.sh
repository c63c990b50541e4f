# usage: bash scripts/gpu_haloed_ab.sh "<variants>"  -- C3 with haloed output, both dissipation forms, per tuning build
for v in $1; do
  if [ "$v" = main ]; then unset EXAHYPE_CUDA_LIB; else export EXAHYPE_CUDA_LIB=$PWD/exahype_b200/variants/$v/libexahype_cuda.so; fi
  for d in var0 all; do
  python bench.py --output haloed --dissipation $d --no-cpu --no-e2e --no-others --no-sustained --steps 30 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$v haloed $d', round(d['roofline']['kernel_ms'],4))"
  done
done
