for v in main noearly r3 r3sb2 main; do
  if [ "$v" = main ]; then unset EXAHYPE_CUDA_LIB; else export EXAHYPE_CUDA_LIB=$PWD/exahype_b200/variants/$v/libexahype_cuda.so; fi
  for o in haloed; do for d in var0 all; do
  python bench.py --output $o --dissipation $d --no-cpu --no-e2e --no-others --no-sustained --steps 30 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$v $o $d', round(d['roofline']['kernel_ms'],4))"
  done; done
done
