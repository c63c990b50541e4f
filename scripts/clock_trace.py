"""How the C3 kernel's per-launch time and the SM clock evolve under sustained back-to-back launches:
python scripts/clock_trace.py [workload] [launches]"""
import os, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from exahype_b200 import runtime

wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
total = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
model, dim, P, h, nr, na, dtype, batch, _ = bench.WORKLOADS[wl]
tdt = torch.float64 if dtype == "f64" else torch.float32
upd = runtime.PatchUpdate(model, dim, P, h, nr, na, dtype=dtype, output="unhaloed")
q = upd.fill_synthetic(torch.empty(upd.in_shape(batch), dtype=tdt, device="cuda"), 0)
out = torch.empty(upd.out_shape(batch), dtype=tdt, device="cuda")
lam = torch.zeros(1, dtype=tdt, device="cuda")
print(subprocess.run(["nvidia-smi", "--query-gpu=power.limit,power.default_limit,power.max_limit,clocks.max.sm,clocks.max.mem",
                      "--format=csv"], capture_output=True, text=True).stdout)
smi = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,power.draw,power.draw.instant,clocks_event_reasons.sw_power_cap",
                        "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE, text=True)
time.sleep(0.5)
upd.step(q, out, 0.01, None, lam)
torch.cuda.synchronize()
time.sleep(0.3)
block = 50
ev = [torch.cuda.Event(enable_timing=True) for _ in range(total // block + 1)]
t0 = time.perf_counter()
ev[0].record()
for i in range(total // block):
    for _ in range(block):
        upd.step(q, out, 0.01, None, lam)
    ev[i + 1].record()
torch.cuda.synchronize()
wall = time.perf_counter() - t0
time.sleep(0.2)
smi.terminate()
lines = smi.communicate()[0].strip().splitlines()
ms = [ev[i].elapsed_time(ev[i + 1]) / block for i in range(len(ev) - 1)]
print(f"{wl}: {total} launches in {wall * 1e3:.0f} ms")
print("ms per launch by block of 50:", " ".join(f"{x:.4f}" for x in ms))
print("nvidia-smi every 20 ms (sm MHz, mem MHz, W avg, W instant, sw_power_cap):")
for l in lines:
    print("  ", l)
