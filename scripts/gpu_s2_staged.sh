# Session-2: unknowns-only output, staged vector stores (main) against one store per value (variant direct_out)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "unknowns" 2>&1 | tail -3
for v in direct_out main; do
  if [ "$v" = main ]; then unset EXAHYPE_CUDA_LIB; else export EXAHYPE_CUDA_LIB=$PWD/exahype_b200/variants/$v/libexahype_cuda.so; fi
  for wl in c4 c4f32 swe_source; do
  python bench.py --workload $wl --output unknowns --no-cpu --no-e2e --no-others --no-fast-leg --steps 20 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); s=d.get('sustained') or {}
print('$v $wl unknowns burst ms %.4f frac %.3f | sustained ms %.4f frac %.3f | bitwise %s' % (d['ms_per_step'], d['roofline']['frac'], s.get('ms_per_step',0), s.get('frac_of_burst_peak',0), d.get('multi_gpu_bitwise')))"
  done
done 2>&1 | tee gpurun_out/s2_staged.txt
