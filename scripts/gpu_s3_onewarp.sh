# Session-3: one-warp CTAs of the warp-per-patch kernel -- 9 per SM (3-deep ring, registers capped at 168 by the
# three-warps-per-scheduler partition) and 8 per SM (4-deep ring, no cap) against the committed 8-warp CTA:
# burst + sustained on C3, then the small-batch end of C5 (4 096 patches)
mkdir -p gpurun_out
for v in $1; do
  if [ "$v" = main ]; then unset EXAHYPE_CUDA_LIB; else export EXAHYPE_CUDA_LIB=$PWD/exahype_b200/variants/$v/libexahype_cuda.so; fi
  python bench.py --workload c3 --no-cpu --no-e2e --no-others --no-fast-leg --steps 30 2>gpurun_out/s3_$v.err | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); s=d.get('sustained') or {}
print('$v c3 burst ms %.4f frac %.3f | sustained ms %.4f frac %.3f | bitwise %s' % (d['ms_per_step'], d['roofline']['frac'], s.get('ms_per_step',0), s.get('frac_of_burst_peak',0), d.get('multi_gpu_bitwise')))"
  python bench.py --workload c3 --batch 4096 --no-cpu --no-e2e --no-others --no-fast-leg --no-sustained --steps 50 2>>gpurun_out/s3_$v.err | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('$v c3 4096 patches ms %.4f frac %.3f | bitwise %s' % (d['ms_per_step'], d['roofline']['frac'], d.get('multi_gpu_bitwise')))"
done 2>&1 | tee gpurun_out/s3_onewarp.txt
