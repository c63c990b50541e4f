set -x
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02_topo.txt 2>&1
# one GPU first: time loop with the exchange stamps
timeout 200 python bench.py --no-cpu --no-e2e --no-others --no-sustained --trace gpurun_out/r02_trace_1_loop > gpurun_out/r02_single_loop.json 2> gpurun_out/r02_single_loop.err; echo rc=$?
python scripts/exchange_attribution.py gpurun_out/r02_trace_1_loop | tee gpurun_out/r02_attribution_1_loop.txt
EXAHYPE_CUDA_LIB=$PWD/exahype_b200/variants/fast/libexahype_cuda.so timeout 200 python bench.py --no-cpu --no-e2e --no-others > gpurun_out/r02_fast_again.json 2>&1; echo rc=$?
bash scripts/gpu_multi.sh 2 r02
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 scripts/pcie_probe.py > gpurun_out/r02_pcie_probe_2gpu.txt 2>&1; echo pcie rc=$?; cat gpurun_out/r02_pcie_probe_2gpu.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu > gpurun_out/r02_bench_2gpu_full.json 2> gpurun_out/r02_bench_2gpu_full.err; echo rc=$?
