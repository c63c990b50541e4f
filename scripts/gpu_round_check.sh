set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/s2_pytest.log 2>&1; echo pytest rc=$?
python bench.py > gpurun_out/s2_bench.json 2> gpurun_out/s2_bench.err; echo bench rc=$?
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/s2_bench_ref.json 2>&1; echo ref rc=$?
python bench.py --no-cpu --no-e2e --variants --steps 10 > gpurun_out/s2_variants.json 2>&1; echo var rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s2_launches.csv python bench.py --no-cpu --steps 5 --warmup 3 > gpurun_out/s2_ncu_list.log 2>&1; echo ncu list rc=$?
ncu --set full --clock-control none --import-source on -k regex:fv3d_march -s 3 -c 1 -o gpurun_out/s2_c3 python bench.py --no-cpu --no-e2e --steps 5 > gpurun_out/s2_ncu_c3.log 2>&1; echo ncu c3 rc=$?
ncu --set full --clock-control none --import-source on -k regex:fv_step -s 3 -c 1 -o gpurun_out/s2_c2 python bench.py --workload c2 --no-cpu --no-e2e --steps 5 > gpurun_out/s2_ncu_c2.log 2>&1; echo ncu c2 rc=$?
ncu --set full --clock-control none --import-source on -k regex:fv_step -s 3 -c 1 -o gpurun_out/s2_c4 python bench.py --workload c4 --no-cpu --no-e2e --steps 5 > gpurun_out/s2_ncu_c4.log 2>&1; echo ncu c4 rc=$?
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv > gpurun_out/s2_smi.txt; nproc >> gpurun_out/s2_smi.txt; lscpu | head -20 >> gpurun_out/s2_smi.txt
