"""What the box's PCIe links give (pinned memory, large copies) and how close exahype_cuda_time_step_host gets.

    python scripts/pcie_probe.py [--sweep]                      one GPU (--sweep: chunk size / depth matrix)
    python -m torch.distributed.run --nproc-per-node N ... scripts/pcie_probe.py      N GPUs copying AT THE SAME TIME

Under torchrun every rank drives its own GPU and all ranks start each measurement together (barrier), so the numbers
are the per-GPU host<->device rates when N links share the host's memory system and PCIe root complexes -- the floor
of bench.py's `e2e` leg at N GPUs.  Rank 0 prints per-rank and aggregate rates.
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from exahype_b200 import runtime

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

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


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def timed(fn, reps=5):
    fn(); barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    t = (time.perf_counter() - t0) / reps
    if world > 1:
        all_t = [None] * world
        dist.all_gather_object(all_t, t)
        return all_t
    return [t]


def h2d():
    with torch.cuda.stream(s1):
        d_in.copy_(h_in, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2):
        h_out.copy_(d_out, non_blocking=True)
def both():
    h2d(); d2h()


def report(name, ts, gbs):
    if rank != 0:
        return
    worst = max(ts)
    parts = "  ".join(f"{name_} {g / worst:.1f}" for name_, g in gbs)
    per_rank = " ".join(f"{gbs[0][1] / t:.1f}" for t in ts)
    print(f"{name:<46s} {worst * 1e3:8.2f} ms (slowest rank)  per GPU GB/s: {parts}   aggregate x{world}: "
          f"{sum(g for _, g in gbs) * world / worst:.1f} GB/s   [{gbs[0][0]} per rank: {per_rank}]", flush=True)


if rank == 0:
    try:
        nodes = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")]
    except OSError:
        nodes = []
    print(f"host: {len(os.sched_getaffinity(0))} cores in the affinity mask, {len(nodes)} NUMA node(s); {world} GPU(s) copying simultaneously; "
          f"{gb_in:.2f} GB in / {gb_out:.2f} GB out per GPU per step", flush=True)
report("H2D alone (pinned, one cudaMemcpyAsync)", timed(h2d), [("H2D", gb_in)])
report("D2H alone", timed(d2h), [("D2H", gb_out)])
report("H2D + D2H concurrently (two streams)", timed(both), [("H2D", gb_in), ("D2H", gb_out)])
lib = runtime.load()
configs = [(0, 0)]
if "--sweep" in sys.argv:
    configs += [(32, 3), (16, 3), (64, 3), (128, 3), (32, 4), (64, 4), (8, 4), (32, 2)]
for chunk_mib, depth in configs:
    lib.exahype_cuda_host_pipeline_release()
    chunk = max(1, (chunk_mib << 20) // (h_in[0].numel() * 8)) if chunk_mib else 0
    lib.exahype_cuda_host_pipeline_configure(chunk, depth)
    ts = timed(lambda: upd.time_step(h_in.numpy(), 0.01, Q_out=h_out.numpy()), reps=3)
    label = f"chunk {chunk_mib} MiB depth {depth}" if chunk_mib else "default pipeline"
    report(f"exahype_cuda_time_step_host ({label})", ts, [("H2D", gb_in), ("D2H", gb_out)])
lib.exahype_cuda_host_pipeline_release()
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
