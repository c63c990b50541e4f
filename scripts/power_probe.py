"""Power / SM clock of a plain device-to-device copy under sustained load, for comparison with the patch kernels
(scripts/clock_trace.py): python scripts/power_probe.py"""
import subprocess, sys, time
import torch

n = 1 << 29   # 512 Mi doubles?  no: bf16-sized elements as in MEASURED_PEAKS: 1 Gi elements of 2 bytes = 2 GiB
a = torch.empty(1 << 30, dtype=torch.bfloat16, device="cuda").normal_()
b = torch.empty_like(a)
smi = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,power.draw,power.draw.instant,clocks_event_reasons.sw_power_cap",
                        "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
time.sleep(0.4)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(41)]
ev[0].record()
for i in range(40):
    for _ in range(10):
        b.copy_(a)
    ev[i + 1].record()
torch.cuda.synchronize()
time.sleep(0.3)
smi.terminate()
lines = smi.communicate()[0].strip().splitlines()
gbs = [2 * a.numel() * 2 * 10 / (ev[i].elapsed_time(ev[i + 1]) * 1e-3) / 1e9 for i in range(40)]
print("copy GB/s (read+write) by block of 10 copies:", " ".join(f"{x:.0f}" for x in gbs))
prev = None
for l in lines:
    if l != prev:
        print("  ", l)
    prev = l
