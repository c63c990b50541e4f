# Session-2: GPU tests, generated-kernel perf (SymPy bodies with the derived primitive cache), and the un-haloed output
# with / without the auxiliary variables on the shallow-water workloads (burst + sustained).
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/s2_pytest.log 2>&1; echo pytest rc=$?; tail -5 gpurun_out/s2_pytest.log
python scripts/generated_kernel_perf.py > gpurun_out/r02_generated_kernel_perf.txt 2>&1; echo gen rc=$?; tail -12 gpurun_out/r02_generated_kernel_perf.txt
set +x
for wl in c4 c4f32 swe_source; do
  for out in unhaloed unknowns; do
  python bench.py --workload $wl --output $out --no-cpu --no-e2e --no-others --no-fast-leg --steps 20 2>gpurun_out/s2_unknowns_$wl_$out.err | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); s=d.get('sustained') or {}
print('$wl $out burst ms %.4f frac %.3f | sustained ms %.4f frac %.3f | bitwise %s' % (d['ms_per_step'], d['roofline']['frac'], s.get('ms_per_step',0), s.get('frac_of_burst_peak',0), d.get('multi_gpu_bitwise')))"
  done
done 2>&1 | tee gpurun_out/s2_unknowns.txt
