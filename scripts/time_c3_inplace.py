"""C3 in the reference's own call shape -- haloed, IN PLACE (q_out == q_in) -- next to the out-of-place forms:
python scripts/time_c3_inplace.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from exahype_b200 import runtime

n = 32768
def timed(fn, reps=30):
    for _ in range(5):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

for diss in ("var0", "all"):
    uh = runtime.PatchUpdate("euler", 3, 8, 1, 5, 0, output="unhaloed", dissipation=diss)
    ha = runtime.PatchUpdate("euler", 3, 8, 1, 5, 0, output="haloed", dissipation=diss)
    q = uh.fill_synthetic(torch.empty(uh.in_shape(n), dtype=torch.float64, device="cuda"), 0)
    out_u = torch.empty(uh.out_shape(n), dtype=torch.float64, device="cuda")
    out_h = torch.empty(ha.out_shape(n), dtype=torch.float64, device="cuda")
    lam = torch.zeros(1, dtype=torch.float64, device="cuda")
    t_u = timed(lambda: uh.step(q, out_u, 0.01, None, lam))
    t_h = timed(lambda: ha.step(q, out_h, 0.01, None, lam))
    # in place: the state is restored from q before every launch (outside the timed events), so it stays the admissible
    # synthetic input
    w = q.clone()
    ha.step(w, None, 0.01, None, lam)
    t_i = 0.0
    for _ in range(20):
        w.copy_(q)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); ha.step(w, None, 0.01, None, lam); b.record()
        torch.cuda.synchronize()
        t_i += a.elapsed_time(b) / 20
    ok = bool(torch.isfinite(w).all())
    print(f"dissipation={diss}: un-haloed out of place {t_u:.4f} ms | haloed out of place {t_h:.4f} ms | haloed in place {t_i:.4f} ms (state finite: {ok})")
