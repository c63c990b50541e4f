"""What the box's PCIe link gives (pinned memory, large copies) and how close exahype_cuda_time_step_host gets for
different chunk sizes / depths: python scripts/pcie_probe.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from exahype_b200 import runtime

n = 32768
upd = runtime.PatchUpdate("euler", 3, 8, 1, 5, 0, output="unhaloed")
h_in = torch.empty(upd.in_shape(n), dtype=torch.float64).pin_memory()
h_out = torch.empty(upd.out_shape(n), dtype=torch.float64).pin_memory()
d_in = torch.empty(upd.in_shape(n), dtype=torch.float64, device="cuda")
d_out = torch.empty(upd.out_shape(n), dtype=torch.float64, device="cuda")
upd.fill_synthetic(d_in, 0)
h_in.copy_(d_in)
gb_in, gb_out = h_in.numel() * 8 / 1e9, h_out.numel() * 8 / 1e9
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps

def h2d():
    with torch.cuda.stream(s1):
        d_in.copy_(h_in, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2):
        h_out.copy_(d_out, non_blocking=True)
def both():
    h2d(); d2h()

t = timed(h2d); print(f"H2D alone   {gb_in:.2f} GB in {t*1e3:.2f} ms = {gb_in/t:.1f} GB/s")
t = timed(d2h); print(f"D2H alone   {gb_out:.2f} GB in {t*1e3:.2f} ms = {gb_out/t:.1f} GB/s")
t = timed(both); print(f"both        {t*1e3:.2f} ms: H2D {gb_in/t:.1f} GB/s, D2H {gb_out/t:.1f} GB/s")
lib = runtime.load()
for chunk_mib, depth in [(32, 3), (16, 3), (64, 3), (128, 3), (32, 4), (64, 4), (8, 4), (32, 2)]:
    chunk = max(1, (chunk_mib << 20) // (h_in[0].numel() * 8))
    lib.exahype_cuda_host_pipeline_release()
    lib.exahype_cuda_host_pipeline_configure(chunk, depth)
    t = timed(lambda: upd.time_step(h_in.numpy(), 0.01, Q_out=h_out.numpy()), reps=3)
    print(f"time_step_host chunk {chunk_mib:4d} MiB ({chunk} patches) depth {depth}: {t*1e3:.2f} ms, H2D {gb_in/t:.1f} GB/s")
