# ncu captures of the 2-D row-marching kernel (C2, C4 fp32) exported as CSV pages on the box
mkdir -p gpurun_out
for w in c2 c4f32; do
python bench.py --workload $w --no-cpu --no-e2e --no-others --no-sustained --steps 5 > gpurun_out/rc_plain_$w.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:fv2d_march -s 5 -c 1 -f -o gpurun_out/rc_$w python bench.py --workload $w --no-cpu --no-e2e --no-others --no-sustained --steps 5 > gpurun_out/rc_ncu_$w.log 2>&1; echo ncu $w rc=$?
done
for r in gpurun_out/rc_*.ncu-rep; do
  b=${r%.ncu-rep}
  ncu -i $r --page raw --csv > $b.raw.csv 2>/dev/null
  ncu -i $r --page source --csv 2>/dev/null | gzip > $b.src.csv.gz
  rm -f $r
done
ls -la gpurun_out
