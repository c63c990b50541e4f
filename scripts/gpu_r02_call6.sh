set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest6.log 2>&1; echo pytest rc=$?; tail -6 gpurun_out/r02_pytest6.log
timeout 400 bash scripts/gpu_arith_variants.sh "main norec main_r3sb2 main" > gpurun_out/r02_c3_variants.txt 2>&1; grep -E "burst|bitwise|differs" gpurun_out/r02_c3_variants.txt
timeout 400 python bench.py > gpurun_out/r02_bench6.json 2> gpurun_out/r02_bench6.err; echo bench rc=$?; tail -3 gpurun_out/r02_bench6.err
python bench.py --workload swe_source --no-cpu --no-e2e --no-others > gpurun_out/r02_bench6_swe_source.json 2>gpurun_out/r02_bench6_swe_source.err; echo rc=$?; tail -2 gpurun_out/r02_bench6_swe_source.err
# ncu: headline kernel, reference and fast arithmetic; launch list of a short bench run
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fv3d_pair -s 5 -c 1 -o gpurun_out/r02_c3_pair python bench.py --no-cpu --no-e2e --no-others --no-sustained --steps 5 > gpurun_out/r02_ncu_c3_pair.log 2>&1; echo ncu rc=$?
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fv3d_pair -s 5 -c 1 -o gpurun_out/r02_c3_pair_fast python bench.py --arithmetic fast --no-cpu --no-e2e --no-others --no-sustained --steps 5 > gpurun_out/r02_ncu_c3_pair_fast.log 2>&1; echo ncu fast rc=$?
python bench.py --no-cpu --no-others --no-sustained --steps 10 > gpurun_out/r02_plain.log 2>&1 && timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --no-cpu --no-others --no-sustained --steps 10 > gpurun_out/r02_ncu_list.log 2>&1; echo ncu list rc=$?
