set -x
mkdir -p gpurun_out
python scripts/generated_kernel_perf.py > gpurun_out/r02_generated_kernel_perf.txt 2>&1; echo gen rc=$?; cat gpurun_out/r02_generated_kernel_perf.txt | tail -14
for v in main whatif2d; do
  if [ "$v" = main ]; then unset EXAHYPE_CUDA_LIB; else export EXAHYPE_CUDA_LIB=$PWD/exahype_b200/variants/$v/libexahype_cuda.so; fi
  for wl in c2 c4 c4f32; do
  python bench.py --workload $wl --no-cpu --no-e2e --no-others --no-fast-leg --steps 20 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); s=d.get('sustained') or {}
print('$v $wl burst ms %.4f frac %.3f | sustained ms %.4f frac %.3f' % (d['ms_per_step'], d['roofline']['frac'], s.get('ms_per_step',0), s.get('frac_of_burst_peak',0)))"
  done
done
unset EXAHYPE_CUDA_LIB
