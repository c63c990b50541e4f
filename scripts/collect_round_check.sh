# Condenses the scratch outputs of scripts/gpu_round_check.sh (gpurun_out/rc_*) into the tracked files under profiles/.
# usage: bash scripts/collect_round_check.sh [round-prefix, default r02]
R=${1:-r02}
G=gpurun_out
P=profiles
last() { tail -1 "$1" > "$2"; }
cp $G/rc_pytest.log $P/${R}_pytest_gpu.log
last $G/rc_bench.json $P/${R}_bench_c3.json
last $G/rc_bench_ref.json $P/${R}_bench_reference_arm.json
last $G/rc_variants.json $P/${R}_bench_variants.json
last $G/rc_bench_swe_source.json $P/${R}_bench_swe_source.json
last $G/rc_bench_c4_unknowns.json $P/${R}_bench_c4_unknowns_only.json
last $G/rc_bench_c2.json $P/${R}_bench_c2.json
python scripts/ncu_summary.py launches $G/rc_launches.csv $P/${R}_launches_bench_c3.txt
python scripts/ncu_summary.py report $G/rc_c3.raw.csv $P/${R}_ncu_c3_pair.txt c3
python scripts/ncu_summary.py report $G/rc_c3_fast.raw.csv $P/${R}_ncu_c3_pair_fast.txt
python scripts/ncu_summary.py report $G/rc_c2.raw.csv $P/${R}_ncu_c2_march2d.txt c2
python scripts/ncu_summary.py report $G/rc_c4.raw.csv $P/${R}_ncu_c4_march2d.txt c4
python scripts/ncu_summary.py report $G/rc_c4f32.raw.csv $P/${R}_ncu_c4f32_march2d.txt c4f32
python scripts/ncu_summary.py report $G/rc_c4_unknowns.raw.csv $P/${R}_ncu_c4_unknowns_only_march2d.txt c4_unknowns
# SASS opcode histograms of the headline kernels (dense, un-haloed, var0) from the built objects
B=exahype_b200/build
{
python scripts/sass_histogram.py $B/inst_euler3d.o 'fv3d_pair_kernel.*ArithIEEE>, exahype::RusanovUpdate, double, 8, 1, 8, 4, false, true, false, 1> >\('
python scripts/sass_histogram.py $B/inst_fast.o 'fv3d_pair_kernel.*ArithFast>, exahype::RusanovUpdate, double, 8, 1, 8, 3, false, true, false, 2> >\('
python scripts/sass_histogram.py $B/inst_euler2d.o 'fv2d_march_kernel.*ArithIEEE>, exahype::RusanovUpdate, double, 16, 1, 1, 16, false, true, 32, 2, false, false> >\('
python scripts/sass_histogram.py $B/inst_swe2d.o 'fv2d_march_kernel.*SwePhysics<3, 1, exahype::ArithIEEE>, exahype::RusanovUpdate, double, 32, 1, 1, 16, false, true, 32, 3, false, (true|false)> >\('
} > $P/${R}_sass_histogram.txt
