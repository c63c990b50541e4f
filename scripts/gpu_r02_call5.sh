set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_time_loop.py tests/test_gpu_fast_arithmetic.py -m gpu -x -q > gpurun_out/r02_pytest5.log 2>&1; echo pytest rc=$?; tail -4 gpurun_out/r02_pytest5.log
for m in "" "--time-step host"; do python bench.py --no-cpu --no-e2e --no-others --no-sustained $m 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('1gpu [$m] ms/step', round(d['ms_per_step'],4), 'bitwise', d.get('multi_gpu_bitwise'))"; done
bash scripts/gpu_multi.sh 2 r02b
bash scripts/gpu_c5_sweep.sh 2 "4096 16384 65536" | tee gpurun_out/r02_c5_sweep_2gpu.txt
bash scripts/gpu_c5_sweep.sh 1 "1024 4096 8192 16384 32768 65536 131072 262144" | tee gpurun_out/r02_c5_sweep_1gpu.txt
