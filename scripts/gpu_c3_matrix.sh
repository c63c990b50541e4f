# usage: bash scripts/gpu_c3_matrix.sh "<variant names or 'main'>"  -- every C3 variant (output form x dissipation x CellData) per tuning build
for v in $1; do
  if [ "$v" = main ]; then unset EXAHYPE_CUDA_LIB; else export EXAHYPE_CUDA_LIB=$PWD/exahype_b200/variants/$v/libexahype_cuda.so; fi
  python bench.py --no-cpu --no-e2e --no-sustained --variants --steps 10 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('$v', 'default', round(d['roofline']['kernel_ms'],4))
for k,x in d['variants'].items():
    if k.startswith('c3') and 'cell-kernel' not in k: print('$v', k[:40], round(x['ms'],4))"
done
