# Session-2: C5 small-batch end on one GPU, CTA-interleaved patch assignment (main) against the CTA-major order (variant)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_time_loop.py tests/test_gpu_fast_arithmetic.py -m gpu -x -q 2>&1 | tail -2
for v in $1; do
  if [ "$v" = main ]; then unset EXAHYPE_CUDA_LIB; else export EXAHYPE_CUDA_LIB=$PWD/exahype_b200/variants/$v/libexahype_cuda.so; fi
  echo "== $v"
  bash scripts/gpu_c5_sweep.sh 1 "$2"
done 2>&1 | tee gpurun_out/s2_c5small.txt
