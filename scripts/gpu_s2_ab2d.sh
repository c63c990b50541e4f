# Session-2 A/B of the 2-D row-marching kernel (run under gpurun): GPU tests first, then burst + sustained per build.
# usage: bash scripts/gpu_s2_ab2d.sh "<variants, 'main' = the product library>" "<workloads>"
set -x
mkdir -p gpurun_out
if [ -z "$SKIP_PYTEST" ]; then python -m pytest tests -m gpu -x -q > gpurun_out/s2_pytest.log 2>&1; echo pytest rc=$?; tail -4 gpurun_out/s2_pytest.log; fi
set +x
for v in $1; do
  if [ "$v" = main ]; then unset EXAHYPE_CUDA_LIB; else export EXAHYPE_CUDA_LIB=$PWD/exahype_b200/variants/$v/libexahype_cuda.so; fi
  for wl in $2; do
  python bench.py --workload $wl --no-cpu --no-e2e --no-others --no-fast-leg --steps 20 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); s=d.get('sustained') or {}
print('$v $wl burst ms %.4f frac %.3f | sustained ms %.4f frac %.3f' % (d['ms_per_step'], d['roofline']['frac'], s.get('ms_per_step',0), s.get('frac_of_burst_peak',0)))"
  done
done 2>&1 | tee gpurun_out/s2_ab2d.txt
unset EXAHYPE_CUDA_LIB
