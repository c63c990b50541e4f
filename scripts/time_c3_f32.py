"""Kernel-only timing of the 3-D Euler 8^3 fp32 instantiation (no BASELINE workload of its own): python scripts/time_c3_f32.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from exahype_b200 import runtime

upd = runtime.PatchUpdate("euler", 3, 8, 1, 5, 0, dtype="f32", output="unhaloed")
n = 32768
q = upd.fill_synthetic(torch.empty(upd.in_shape(n), dtype=torch.float32, device="cuda"), 0)
out = torch.empty(upd.out_shape(n), dtype=torch.float32, device="cuda")
lam = torch.zeros(1, dtype=torch.float32, device="cuda")
for _ in range(20):
    upd.step(q, out, 0.01, None, lam)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); a.record()
for _ in range(50):
    upd.step(q, out, 0.01, None, lam)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 50
print(f"c3 f32: {ms:.4f} ms, {upd.algorithmic_bytes_per_patch * n / ms / 1e6:.0f} GB/s algorithmic, launch {upd.launch_info(n)}")
