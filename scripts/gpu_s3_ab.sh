# Session-3 A/B of tuning builds of the warp-per-patch kernel inside one call: 3-D parity tests with the variant library,
# then burst + sustained on C3 and the 4 096-patch end of C5, main and variant alternating twice
# usage: bash scripts/gpu_s3_ab.sh "<variant names>"
mkdir -p gpurun_out
for v in $1; do
  EXAHYPE_CUDA_LIB=$PWD/exahype_b200/variants/$v/libexahype_cuda.so python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "3 or c3 or cell_data or extreme" 2>&1 | tail -1
done
for rep in 1 2; do
for v in main $1; do
  if [ "$v" = main ]; then unset EXAHYPE_CUDA_LIB; else export EXAHYPE_CUDA_LIB=$PWD/exahype_b200/variants/$v/libexahype_cuda.so; fi
  python bench.py --workload c3 --no-cpu --no-e2e --no-others --no-fast-leg --steps 30 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); s=d.get('sustained') or {}
print('$v c3 burst ms %.4f frac %.3f | sustained ms %.4f frac %.3f | bitwise %s' % (d['ms_per_step'], d['roofline']['frac'], s.get('ms_per_step',0), s.get('frac_of_burst_peak',0), d.get('multi_gpu_bitwise')))"
  python bench.py --workload c3 --batch 4096 --no-cpu --no-e2e --no-others --no-fast-leg --no-sustained --steps 50 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('$v c3 4096 patches ms %.4f frac %.3f | bitwise %s' % (d['ms_per_step'], d['roofline']['frac'], d.get('multi_gpu_bitwise')))"
done
done 2>&1 | tee gpurun_out/s3_ab.txt
