set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/r02_smi.txt; nproc >> gpurun_out/r02_smi.txt
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest1.log 2>&1; echo pytest rc=$?; tail -15 gpurun_out/r02_pytest1.log
timeout 400 python bench.py > gpurun_out/r02_bench1.json 2> gpurun_out/r02_bench1.err; echo bench rc=$?; tail -3 gpurun_out/r02_bench1.err
timeout 200 python bench.py --time-step host --no-cpu --no-e2e --no-others > gpurun_out/r02_bench1_hostdt.json 2> gpurun_out/r02_bench1_hostdt.err; echo bench-host rc=$?
timeout 400 bash scripts/gpu_arith_variants.sh "main fma fast" > gpurun_out/r02_arith.txt 2>&1; cat gpurun_out/r02_arith.txt
