# Round check on one B200 (run under gpurun): GPU tests, smoke, the default bench line, the reference arm, all workload
# variants, the launch list and one `--set full` capture per headline kernel.  Outputs land in gpurun_out/rc_*.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/rc_pytest.log 2>&1; echo pytest rc=$?; tail -4 gpurun_out/rc_pytest.log
python __graft_entry__.py smoke > gpurun_out/rc_smoke.log 2>&1; echo smoke rc=$?; tail -1 gpurun_out/rc_smoke.log
python bench.py > gpurun_out/rc_bench.json 2> gpurun_out/rc_bench.err; echo bench rc=$?
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/rc_bench_ref.json 2>&1; echo ref rc=$?
python bench.py --no-cpu --no-e2e --no-sustained --variants --steps 10 > gpurun_out/rc_variants.json 2>&1; echo var rc=$?
python bench.py --workload swe_source --no-cpu --no-e2e --no-others > gpurun_out/rc_bench_swe_source.json 2> gpurun_out/rc_bench_swe_source.err; echo swe_source rc=$?
python bench.py --no-cpu --no-others --no-sustained --steps 10 > gpurun_out/rc_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/rc_launches.csv python bench.py --no-cpu --no-others --no-sustained --steps 10 > gpurun_out/rc_ncu_list.log 2>&1; echo ncu list rc=$?
python bench.py --no-cpu --no-e2e --no-others --no-sustained --steps 5 > gpurun_out/rc_plain_c3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:fv3d_pair -s 5 -c 1 -f -o gpurun_out/rc_c3 python bench.py --no-cpu --no-e2e --no-others --no-sustained --steps 5 > gpurun_out/rc_ncu_c3.log 2>&1; echo ncu c3 rc=$?
python bench.py --arithmetic fast --no-cpu --no-e2e --no-others --no-sustained --steps 5 > gpurun_out/rc_plain_c3_fast.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:fv3d_pair -s 5 -c 1 -f -o gpurun_out/rc_c3_fast python bench.py --arithmetic fast --no-cpu --no-e2e --no-others --no-sustained --steps 5 > gpurun_out/rc_ncu_c3_fast.log 2>&1; echo ncu c3 fast rc=$?
for w in c2 c4 c4f32; do
python bench.py --workload $w --no-cpu --no-e2e --no-others --no-sustained --steps 5 > gpurun_out/rc_plain_$w.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:fv2d_march -s 5 -c 1 -f -o gpurun_out/rc_$w python bench.py --workload $w --no-cpu --no-e2e --no-others --no-sustained --steps 5 > gpurun_out/rc_ncu_$w.log 2>&1; echo ncu $w rc=$?
done
# C4 with the un-haloed output that does not repeat the auxiliary variable (EXAHYPE_FLAG_OUTPUT_UNKNOWNS_ONLY)
python bench.py --workload c4 --output unknowns --no-cpu --no-e2e --no-others --steps 20 > gpurun_out/rc_bench_c4_unknowns.json 2> gpurun_out/rc_bench_c4_unknowns.err; echo c4 unknowns rc=$?
ncu --set full --clock-control none --import-source on -k regex:fv2d_march -s 5 -c 1 -f -o gpurun_out/rc_c4_unknowns python bench.py --workload c4 --output unknowns --no-cpu --no-e2e --no-others --no-sustained --steps 5 > gpurun_out/rc_ncu_c4_unknowns.log 2>&1; echo ncu c4 unknowns rc=$?
python bench.py --workload c2 --no-cpu --no-e2e --no-others --steps 20 > gpurun_out/rc_bench_c2.json 2> gpurun_out/rc_bench_c2.err; echo c2 rc=$?
# gpurun copies back at most 64 MiB: keep the pages as CSV, the reports themselves only for the two headline kernels
for r in gpurun_out/rc_*.ncu-rep; do
  b=${r%.ncu-rep}
  ncu -i $r --page raw --csv > $b.raw.csv 2>/dev/null
  ncu -i $r --page source --csv 2>/dev/null | gzip > $b.src.csv.gz
  case $b in *rc_c3|*rc_c2) ;; *) rm -f $r ;; esac
done
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv > gpurun_out/rc_smi.txt; nproc >> gpurun_out/rc_smi.txt
