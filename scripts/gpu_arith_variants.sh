# usage: bash scripts/gpu_arith_variants.sh "<variant names or 'main'>"  -- per tuning build (exahype_b200/variants/<name>):
# accuracy against the oracle on C3-shaped input, then burst and sustained C3 timings
for v in $1; do
  if [ "$v" = main ]; then unset EXAHYPE_CUDA_LIB; else export EXAHYPE_CUDA_LIB=$PWD/exahype_b200/variants/$v/libexahype_cuda.so; fi
  python - <<PY
import numpy as np, torch
import oracle as O
from exahype_b200 import runtime
for wl,(model,dim,P,nr,na) in {"c3":("euler",3,8,5,0),"c2":("euler",2,16,4,0),"c4":("swe",2,32,3,1)}.items():
    upd = runtime.PatchUpdate(model, dim, P, 1, nr, na, output="haloed")
    cfg = O.OracleConfig(dim=dim, patch_size=P, halo=1, n_real=nr, n_aux=na, model=O.MODEL_EULER if model=="euler" else O.MODEL_SWE)
    q0 = O.fill_synthetic(cfg, 256); want = q0.copy(); lam_o, lmax_o = O.step(cfg, want, 0.01, nthreads=4)
    q = torch.from_numpy(q0).cuda(); lam = torch.zeros(256, dtype=torch.float64, device="cuda")
    upd.step(q, None, 0.01, lam, None); torch.cuda.synchronize()
    got = q.cpu().numpy()
    rel = np.abs(got - want).max() / np.abs(want).max()
    print("$v", wl, "bitwise" if np.array_equal(got, want) else "differs", "max |err| / max |q| = %.3e" % rel,
          "lambda rel err %.3e" % (np.abs(lam.cpu().numpy() - lam_o) / lam_o).max())
PY
  for wl in c3; do
  python bench.py --workload $wl --no-cpu --no-e2e --no-others --steps 20 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())
s=d.get('sustained') or {}; f=d.get('fast_arithmetic') or {}
print('$v $wl burst kernel_ms %.4f frac %.3f | sustained ms %.4f frac %.3f clocks %s' % (d['roofline']['kernel_ms'], d['roofline']['frac'], s.get('ms_per_step',0), s.get('frac_of_burst_peak',0), (s.get('clocks') or {}).get('sm_mhz')))
if f: print('$v $wl FAST  burst ms %.4f frac %.3f | sustained ms %.4f frac %.3f clocks %s' % (f['ms_per_step'], f['frac'], f['sustained']['ms_per_step'], f['sustained']['frac_of_burst_peak'], f['sustained']['clocks'].get('sm_mhz')))"
  done
done
