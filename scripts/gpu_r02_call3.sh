set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest3.log 2>&1; echo pytest rc=$?; tail -8 gpurun_out/r02_pytest3.log
for i in 1 2; do python bench.py --no-cpu --no-e2e --no-others 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); s=d['sustained']
print('main run $i burst', round(d['roofline']['kernel_ms'],4), 'ms/step', round(d['ms_per_step'],4), 'sustained', round(s['ms_per_step'],4), s['clocks']['sm_mhz'], 'bitwise', d['multi_gpu_bitwise'])"; done
timeout 400 bash scripts/gpu_arith_variants.sh "fast fast_r3sb2" > gpurun_out/r02_arith3.txt 2>&1; grep burst gpurun_out/r02_arith3.txt
# ncu: main and fast, one launch of the C3 kernel each
for v in main fast; do
  if [ "$v" = main ]; then unset EXAHYPE_CUDA_LIB; else export EXAHYPE_CUDA_LIB=$PWD/exahype_b200/variants/$v/libexahype_cuda.so; fi
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:fv3d_pair -s 5 -c 1 -o gpurun_out/r02_c3_$v python bench.py --no-cpu --no-e2e --no-others --no-sustained --steps 5 > gpurun_out/r02_ncu_c3_$v.log 2>&1; echo ncu $v rc=$?
done
unset EXAHYPE_CUDA_LIB
