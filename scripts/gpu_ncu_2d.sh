for w in c2 c4; do
python bench.py --workload $w --no-cpu --no-e2e --steps 5 > gpurun_out/s3_plain_$w.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:fv2d_march -s 3 -c 1 -o gpurun_out/s3_$w python bench.py --workload $w --no-cpu --no-e2e --steps 5 > gpurun_out/s3_ncu_$w.log 2>&1; echo ncu $w rc=$?
done
