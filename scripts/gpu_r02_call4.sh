set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest4.log 2>&1; echo pytest rc=$?; tail -8 gpurun_out/r02_pytest4.log
timeout 400 python bench.py > gpurun_out/r02_bench4.json 2> gpurun_out/r02_bench4.err; echo bench rc=$?; tail -3 gpurun_out/r02_bench4.err
show() { python -c "
import sys,json
d=json.loads([l for l in open('$1') if l.startswith('{')][0]); s=d.get('sustained') or {}; f=d.get('fast_arithmetic') or {}
print('$2', 'ms/step', round(d['ms_per_step'],4), 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'frac', round(d['roofline']['frac'],4), 'sustained', round(s.get('ms_per_step',0),4), round(s.get('frac_of_burst_peak',0),3), (s.get('clocks') or {}).get('sm_mhz'), 'bitwise', d.get('multi_gpu_bitwise'))
if f: print('   fast leg: ms/step', round(f['ms_per_step'],4), 'frac', round(f['frac'],4), 'err', f['max_abs_err_over_max_abs_q_vs_reference_arithmetic'], 'sustained', round(f['sustained']['ms_per_step'],4), round(f['sustained']['frac_of_burst_peak'],3), f['sustained']['clocks'].get('sm_mhz'))
"; }
show gpurun_out/r02_bench4.json default
EXAHYPE_PDL=0 python bench.py --no-cpu --no-e2e --no-others --no-sustained > gpurun_out/r02_b4_nopdl.json 2>/dev/null; show gpurun_out/r02_b4_nopdl.json no-pdl
python bench.py --no-cpu --no-e2e --no-others --no-sustained > gpurun_out/r02_b4_pdl.json 2>/dev/null; show gpurun_out/r02_b4_pdl.json pdl
python bench.py --no-cpu --no-e2e --no-others --no-sustained --step-events > gpurun_out/r02_b4_events.json 2>/dev/null; show gpurun_out/r02_b4_events.json step-events
python bench.py --no-cpu --no-e2e --no-others --no-sustained --time-step host > gpurun_out/r02_b4_host.json 2>/dev/null; show gpurun_out/r02_b4_host.json host-dt
for wl in c2 c4 c4f32; do for ar in reference fast; do
python bench.py --workload $wl --arithmetic $ar --no-cpu --no-e2e --no-others > gpurun_out/r02_b4_${wl}_$ar.json 2>/dev/null; show gpurun_out/r02_b4_${wl}_$ar.json "$wl $ar"
done; done
python bench.py --arithmetic fast --no-cpu --no-e2e --no-others > gpurun_out/r02_b4_c3_fast.json 2>/dev/null; show gpurun_out/r02_b4_c3_fast.json "c3 fast"
