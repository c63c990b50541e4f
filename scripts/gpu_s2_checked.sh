# Session-2: the checked default arithmetic (nvcc's fast path + one range check per cell) against the per-operation IEEE
# build: full GPU tests with the new library, then burst + sustained per build on C3, C2, C4
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/s2_pytest.log 2>&1; echo pytest rc=$?; tail -4 gpurun_out/s2_pytest.log
for v in $1; do
  if [ "$v" = main ]; then unset EXAHYPE_CUDA_LIB; else export EXAHYPE_CUDA_LIB=$PWD/exahype_b200/variants/$v/libexahype_cuda.so; fi
  for wl in $2; do
  python bench.py --workload $wl --no-cpu --no-e2e --no-others --no-fast-leg --steps 30 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); s=d.get('sustained') or {}
print('$v $wl burst ms %.4f frac %.3f | sustained ms %.4f frac %.3f | bitwise %s' % (d['ms_per_step'], d['roofline']['frac'], s.get('ms_per_step',0), s.get('frac_of_burst_peak',0), d.get('multi_gpu_bitwise')))"
  done
done 2>&1 | tee gpurun_out/s2_checked.txt
