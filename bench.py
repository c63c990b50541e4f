#!/usr/bin/env python
"""Benchmark of the batched stateless FV Rusanov patch update (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c3|c2|c4|c4f32|c1]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One "step" = one pass of the hot path over the whole batch: fused patch-update kernel (+ the NCCL all-reduce(max) of
the admissible-time-step scalar when N > 1).  Prints ONE JSON line on rank 0.

  value     patch-cell updates/s, whole job, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e       same metric through the reference-facing host call exahype_cuda_time_step_host(Q_host, dt): pinned host
            buffers, H2D and D2H inside the timed region
  roofline  algorithmic bytes / average kernel duration vs the measured HBM copy peak
  cpu_baseline  the CPU oracle (port of the reference kernel's arithmetic) on this box's host cores, bounded sample

Weak scaling: every rank owns `batch` patches (contiguous shard of the global batch, counter-based synthetic input).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (model, dim, P, halo, n_real, n_aux, dtype, patches per GPU, description)
    "c3": ("euler", 3, 8, 1, 5, 0, "f64", 32768, "3D Euler (5 unknowns) Rusanov FV, 8x8x8 patches + 1 halo, batch of 32,768 patches, fp64"),
    "c2": ("euler", 2, 16, 1, 4, 0, "f64", 65536, "2D Euler Rusanov FV, 16x16 patches + 1 halo, batch of 65,536 patches, fp64"),
    "c4": ("swe", 2, 32, 1, 3, 1, "f64", 65536, "2D shallow-water (3 unknowns + bathymetry) Rusanov FV, 32x32 patches, batch of 65,536, fp64"),
    "c4f32": ("swe", 2, 32, 1, 3, 1, "f32", 65536, "2D shallow-water Rusanov FV, 32x32 patches, batch of 65,536, fp32"),
    # not a BASELINE config: the source-term model of SURVEY.md 8f-3 on C4's shape, for its roofline line
    "swe_source": ("swe_source", 2, 32, 1, 3, 3, "f64", 49152, "2D shallow water with bathymetry source term (3 unknowns + b, db/dx, db/dy), 32x32 patches, batch of 49,152, fp64"),
    "c1": ("euler", 2, 3, 1, 4, 0, "f64", 1000, "2D Euler (4 unknowns) Rusanov FV, 3x3 patches + 1 halo, batch of 1,000 patches, fp64"),
}
METRIC = "patch_cell_updates_per_sec"
UNIT = "cell-updates/s"


def measured_hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def recorded_traffic(workload: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu capture, if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get(workload)
    except Exception:
        return None


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, flag in zip(names, f[5:9]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------------ CPU arm
def host_threads() -> int:
    """All the host cores this process may run on.  Not omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1 to its
    workers, which would time the CPU arm on one core."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_reference_rate(workload: str, min_seconds: float, sample_patches: int):
    """Times the CPU oracle (contraction-free -O3 build, OpenMP over patches: all host threads) on a bounded sample of
    the workload.  Returns (cell-updates/s, description dict)."""
    import numpy as np
    import oracle as O
    model, dim, P, h, nr, na, dtype, _, _ = WORKLOADS[workload]
    cfg = O.OracleConfig(dim=dim, patch_size=P, halo=h, n_real=nr, n_aux=na,
                         model={"euler": O.MODEL_EULER, "swe": O.MODEL_SWE, "swe_source": O.MODEL_SWE_SOURCE}[model])
    threads = host_threads()
    npdt = np.float64 if dtype == "f64" else np.float32
    q0 = O.fill_synthetic(cfg, sample_patches, dtype=npdt)
    q = q0.copy()
    O.step(cfg, q, 0.01, nthreads=threads, fast=True)          # warm-up (page faults, thread pool)
    passes, elapsed, t0 = 0, 0.0, time.perf_counter()
    while time.perf_counter() - t0 < min_seconds or passes < 3:
        np.copyto(q, q0)                                        # restore outside the timed region: the state stays admissible
        t1 = time.perf_counter()
        O.step(cfg, q, 0.01, nthreads=threads, fast=True)
        elapsed += time.perf_counter() - t1
        passes += 1
    cells = sample_patches * P ** dim * passes
    return cells / elapsed, {"cores": threads, "kind": "port",
                             "sample": f"{sample_patches} patches of the workload x {passes} passes "
                                       f"({elapsed:.1f} s), OpenMP over patches, gcc -O3 -march=native -ffp-contract=off"}


def run_reference_arm(args):
    """`--impl reference`: the reference kernel's arithmetic on the host CPU (the oracle port; the reference's own
    compiled kernel, oracle/_ref, is hard-wired to one 4x4 2-D patch and cannot run this workload)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    model, dim, P, h, nr, na, dtype, batch, desc = WORKLOADS[args.workload]
    import numpy as np
    import oracle as O
    cfg = O.OracleConfig(dim=dim, patch_size=P, halo=h, n_real=nr, n_aux=na,
                         model={"euler": O.MODEL_EULER, "swe": O.MODEL_SWE, "swe_source": O.MODEL_SWE_SOURCE}[model])
    threads = host_threads()
    sample = min(batch, args.cpu_sample)
    npdt = np.float64 if dtype == "f64" else np.float32
    q0 = O.fill_synthetic(cfg, sample, dtype=npdt)
    q = q0.copy()
    for _ in range(max(1, args.warmup)):
        np.copyto(q, q0)
        O.step(cfg, q, 0.01, nthreads=threads, fast=True)
    elapsed = 0.0
    for _ in range(args.steps):
        np.copyto(q, q0)                                        # restore (untimed) so every step sees admissible data
        t0 = time.perf_counter()
        O.step(cfg, q, 0.01, nthreads=threads, fast=True)
        elapsed += time.perf_counter() - t0
    value = sample * P ** dim * args.steps / elapsed
    ref = O.reference_compiled_rate(2.0)      # the reference's own compiled kernel on its hard-wired 4x4 shape, 1 thread
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * elapsed / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": dtype, "data": "synthetic",
            "config": {"workload": desc, "patches_per_step": sample, "note": "bounded sample of the workload per step"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"{sample} patches per step x {args.steps} steps, OpenMP over patches",
                             **({"reference_compiled": {"value": ref[0], "unit": UNIT, "cores": 1, "sample": ref[1]}}
                                if ref is not None else {})},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def bind_to_gpu_numa_node(torch, device_index: int):
    """Best effort: run this process on the cores of the NUMA node its GPU hangs off, so that the pinned host buffers of
    the e2e leg (first touch) are local to the PCIe root the copies go through.  Returns the node or None."""
    try:
        bus = torch.cuda.get_device_properties(device_index).pci_bus_id          # e.g. 0000:1B:00.0 (torch >= 2.3)
    except Exception:
        try:
            import ctypes
            buf = ctypes.create_string_buffer(32)
            rt = ctypes.CDLL("libcudart.so")
            if rt.cudaDeviceGetPCIBusId(buf, 32, device_index) != 0:
                return None
            bus = buf.value.decode()
        except Exception:
            return None
    try:
        if isinstance(bus, int):
            return None
        path = f"/sys/bus/pci/devices/{bus.lower()}/numa_node"
        node = int(open(path).read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="patches per GPU (default: the workload's BASELINE batch)")
    ap.add_argument("--output", default="unhaloed", choices=["unhaloed", "haloed", "unknowns"],
                    help="unknowns: the un-haloed output without the auxiliary variables (EXAHYPE_FLAG_OUTPUT_UNKNOWNS_ONLY)")
    ap.add_argument("--dissipation", default="var0", choices=["var0", "all"])
    ap.add_argument("--reducer", default="auto", choices=["auto", "peer", "nccl"],
                    help="N > 1: all-reduce(max) of lambda over NVLink peer memory (one-shot kernel) or through NCCL")
    ap.add_argument("--no-fused", action="store_true",
                    help="N > 1: all-reduce as a separate launch instead of the patch kernel's epilogue")
    ap.add_argument("--kernel", default="auto", choices=["auto", "cell"], help="3-D: plane-marching (auto) or thread-per-cell")
    ap.add_argument("--arithmetic", default="reference", choices=["reference", "fast"],
                    help="reference: the reference's arithmetic bit for bit (default, the headline); fast: "
                         "EXAHYPE_FLAG_FAST_ARITHMETIC (contracted FMAs, branch-free 1/x and sqrt; within 1e-12, not bitwise)")
    ap.add_argument("--no-fast-leg", action="store_true", help="skip the informative fast-arithmetic leg of the default run")
    ap.add_argument("--time-step", default="device", choices=["device", "host"],
                    help="device: dt of step k+1 = cfl_dx / global lambda_max of step k, produced and consumed on the device "
                         "(exahype_cuda_fv_step_time_loop); host: dt is a host constant and lambda_max is only reduced")
    ap.add_argument("--cfl-dx", type=float, default=0.4 / 8, help="CFL number x cell size of the device-resident time loop")
    ap.add_argument("--step-events", action="store_true",
                    help="bracket every launch of the timed region with its own CUDA events (roofline.kernel_ms = mean launch "
                         "duration); default: two events around the K steps, so that consecutive launches can overlap their "
                         "start-up with the previous step's tail (programmatic dependent launch), kernel_ms = ms_per_step")
    ap.add_argument("--trace", default="", help="write the exchange's device-side globaltimer stamps of the timed steps to this file (per rank)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--cpu-sample", type=int, default=4096)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-numa-bind", action="store_true", help="e2e leg: do not pin the process to the GPU's NUMA node")
    ap.add_argument("--no-sustained", action="store_true", help="skip the sustained-load (power-limited) leg")
    ap.add_argument("--no-others", action="store_true", help="skip the kernel-only timings of the other BASELINE workloads")
    ap.add_argument("--variants", action="store_true", help="also time the other output/dissipation variants and workloads")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        run_reference_arm(args)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from exahype_b200 import runtime
    from exahype_b200.dist import PatchSharding, TimestepReducer, TimeLoop

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if args.gpus != world and rank == 0:
        print(f"note: --gpus {args.gpus} but WORLD_SIZE={world}; using {world}", file=sys.stderr)

    model, dim, P, h, nr, na, dtype, batch, desc = WORKLOADS[args.workload]
    batch = args.batch or batch
    upd = runtime.PatchUpdate(model, dim, P, h, nr, na, dtype=dtype, dissipation=args.dissipation, output=args.output,
                              kernel=args.kernel, arithmetic=args.arithmetic)
    tdt = torch.float64 if dtype == "f64" else torch.float32
    npdt = np.float64 if dtype == "f64" else np.float32
    shard = PatchSharding(global_patches=batch * world, world_size=world, rank=rank)
    assert shard.count == batch

    # synthetic admissible input of SURVEY.md 8d, generated on the device shard by shard (same bits as the oracle's)
    q_in = synthetic_on_device(torch, upd, shard.first, batch, tdt)
    q_out = torch.empty(upd.out_shape(batch), dtype=tdt, device="cuda")
    lam_patch = torch.empty(batch, dtype=tdt, device="cuda")
    lam_max = torch.zeros(1, dtype=tdt, device="cuda")
    reducer = TimestepReducer(world, rank, backend=args.reducer) if world > 1 else None
    stream = torch.cuda.current_stream()

    # Device-resident time loop (default): step k+1 runs with dt = cfl_dx / max over ranks of lambda_max(step k); the
    # exchange is split-phase over NVLink peer memory (published by step k's last warp, consumed by step k+1's warps),
    # nothing crosses to the host between steps.  It needs the peer-memory reducer at N > 1; with --reducer nccl (or
    # --time-step host) dt is a host constant and lambda_max is only reduced -- the round-1 loop.
    peer_ok = reducer is None or reducer.backend == "peer"
    device_dt = args.time_step == "device" and peer_ok
    tracer = reducer
    if args.trace and world == 1 and device_dt:
        from exahype_b200.dist import LocalPeerGroup
        tracer = LocalPeerGroup(1)[0]            # one GPU: the loop's own one-rank exchange, with stamps
    loop = TimeLoop(dtype, tracer if world == 1 else reducer, args.cfl_dx, 0.01) if device_dt else None
    tracing = bool(args.trace) and tracer is not None and tracer.backend == "peer"
    if tracing:
        tracer.enable_trace(max(64, 2 * (args.steps + args.warmup) + 16))
    # host-dt mode: all-reduce in the patch kernel's own epilogue (blocking form, peer-memory backend)
    fused = (not device_dt) and reducer is not None and reducer.backend == "peer" and not args.no_fused

    def step():
        if device_dt:
            upd.step_loop(loop, q_in, q_out, lam_patch)
        elif fused:
            upd.step(q_in, q_out, 0.01, lam_patch, lam_max, reducer=reducer)
        else:
            upd.step(q_in, q_out, 0.01, lam_patch, lam_max)
            if reducer is not None:
                reducer.allreduce_max(lam_max)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # The clock sampler starts BEFORE the first warm-up step (nvidia-smi needs a few hundred ms to its first sample), so
    # that warm-up and timed steps follow each other with nothing but the contract's barrier in between.  This is a
    # burst measurement like MEASURED_PEAKS.json's copy peak (best of 10): after ~35 ms of back-to-back launches the
    # board reaches its 1000 W power limit and lowers the SM clock (scripts/clock_trace.py, profiles/); the
    # `sustained` leg below reports that regime separately.
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.5)
    barrier()
    for _ in range(args.warmup):
        step()
    launches0 = runtime.launch_count()
    k_start = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    k_stop = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_begin.record(stream)
    for i in range(args.steps):
        if args.step_events:
            k_start[i].record(stream)
        if device_dt:
            upd.step_loop(loop, q_in, q_out, lam_patch)
        elif fused:
            upd.step(q_in, q_out, 0.01, lam_patch, lam_max, reducer=reducer)
        else:
            upd.step(q_in, q_out, 0.01, lam_patch, lam_max)
        if args.step_events:
            k_stop[i].record(stream)
        if not device_dt and not fused and reducer is not None:
            reducer.allreduce_max(lam_max)
    t_end.record(stream)
    barrier()
    launches = runtime.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    if tracing:
        first_seq = loop.steps - args.steps + 1 if device_dt else args.warmup + 1
        np.save(f"{args.trace}.rank{rank}.npy", tracer.read_trace(max(1, first_seq), args.steps))

    # --- correctness of the multi-GPU path, outside the timed region
    if loop is not None:
        loop.flush()                              # the exchange still in flight belongs to the reducer, not to the loop
        torch.cuda.synchronize()
        hist = loop.history(loop.steps - 1, 2)
        lam_max.fill_(float(hist[1, 1]))          # the last step's global maximum
    checks = verify_time_loop(torch, dist, upd, reducer, args, q_in, q_out, lam_patch, shard, world, rank) if device_dt else None
    if reducer is not None and not device_dt:
        # the collective is exact: the reduced scalar must equal torch.distributed's own MAX over the ranks' local values
        upd.step(q_in, q_out, 0.01, lam_patch, lam_max)
        local = lam_max.clone()
        if fused:     # the same step again, reduced by its own epilogue
            upd.step(q_in, q_out, 0.01, lam_patch, lam_max, reducer=reducer)
        else:
            reducer.allreduce_max(lam_max)
        dist.all_reduce(local, op=dist.ReduceOp.MAX)
        torch.cuda.synchronize()
        if float(local.item()) != float(lam_max.item()) or reducer.timed_out():
            raise SystemExit(f"rank {rank}: all-reduce(max) mismatch: {float(lam_max.item())} vs {float(local.item())}")

    elapsed_ms = t_begin.elapsed_time(t_end)
    # per-launch duration: the mean over the launches' own event pairs, or -- without events between the launches -- the
    # step time itself (an upper bound: it contains the launch gaps)
    kernel_ms = (statistics.mean(a.elapsed_time(b) for a, b in zip(k_start, k_stop)) if args.step_events
                 else elapsed_ms / args.steps)
    if world > 1:
        t = torch.tensor([elapsed_ms, kernel_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms, kernel_ms = t.tolist()
    cells_per_step = batch * world * P ** dim
    value = cells_per_step * args.steps / (elapsed_ms * 1e-3)
    lam_global = float(lam_max.item())

    # --- end to end through the host-facing C-ABI call (pinned host buffers, copies inside the timed region)
    e2e = None
    if not args.no_e2e:
        affinity = os.sched_getaffinity(0) if hasattr(os, "sched_getaffinity") else None
        numa_node = None if args.no_numa_bind else bind_to_gpu_numa_node(torch, local_rank)
        host_in = torch.empty(upd.in_shape(batch), dtype=tdt).pin_memory()
        host_in.copy_(q_in)
        host_out = torch.empty(upd.out_shape(batch), dtype=tdt).pin_memory()
        upd.time_step(host_in.numpy(), 0.01, Q_out=host_out.numpy())            # warm-up: allocates the staging ring
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            lam_e2e = upd.time_step(host_in.numpy(), 0.01, Q_out=host_out.numpy())
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = float(t.item())
        # the floor of this leg on this box: plain pinned copies of the same bytes, both directions at once, on every GPU
        # at the same time (no kernel, no chunking) -- what the host's memory system and the PCIe links give N GPUs
        d_probe = torch.empty_like(q_out)
        sa, sb = torch.cuda.Stream(), torch.cuda.Stream()

        def plain_copies():
            with torch.cuda.stream(sa):
                q_in.copy_(host_in, non_blocking=True)
            with torch.cuda.stream(sb):
                host_out.copy_(d_probe, non_blocking=True)
        plain_copies()
        barrier()
        t0 = time.perf_counter()
        for _ in range(2):
            plain_copies()
        torch.cuda.synchronize()
        floor_s = (time.perf_counter() - t0) / 2
        if world > 1:
            t = torch.tensor([floor_s], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            floor_s = float(t.item())
        del d_probe
        q_in = synthetic_on_device(torch, upd, shard.first, batch, tdt)      # the probe overwrote it with the same bits; regenerate anyway
        es = 8 if dtype == "f64" else 4
        e2e = {"value": cells_per_step * args.e2e_steps / e2e_s, "unit": UNIT,
               "h2d_bytes_per_step": int(host_in.numel() * es) * world,
               "d2h_bytes_per_step": int(host_out.numel() * es + es) * world,
               "ms_per_step": 1e3 * e2e_s / args.e2e_steps, "steps": args.e2e_steps,
               "api": "exahype_cuda_time_step_host (chunked H2D -> kernel -> D2H over 3 stream slots)",
               "lambda_max_matches_device": bool(float(lam_e2e) == float(lam_patch.max().item())),
               "host_numa_node": numa_node,
               "h2d_GBs_per_gpu": host_in.numel() * es / (e2e_s / args.e2e_steps) / 1e9,
               "plain_copy_floor_ms": 1e3 * floor_s,
               "floor_over_e2e": floor_s / (e2e_s / args.e2e_steps),
               "floor": "plain pinned cudaMemcpyAsync of the same bytes, H2D and D2H concurrently, all GPUs at once (max over ranks)"}
        del host_in, host_out
        runtime.load().exahype_cuda_host_pipeline_release()
        if numa_node is not None and affinity:
            os.sched_setaffinity(0, affinity)      # the CPU baseline below uses every core again

    variants = None
    if args.variants and world == 1:
        variants = time_variants(torch, runtime, args)
    # the other BASELINE configurations next to the headline one (kernel-only, a few ms each): parity-test cases, not
    # bench lines, reported so that one run shows every committed kernel family against the same roofline
    others = None
    if world == 1 and not args.no_others:
        peak_gbs, _ = measured_hbm_peak()
        others = {}
        for wl in ("c2", "c4", "c4f32"):
            if wl == args.workload:
                continue
            r = time_kernel_only(torch, runtime, wl)
            others[wl] = {"workload": WORKLOADS[wl][8], "kernel_ms": r["ms"], "value": r["cell_updates_per_s"],
                          "achieved_GBs": r["algorithmic_GBs"], "frac": r["algorithmic_GBs"] / peak_gbs}
            if WORKLOADS[wl][5] > 0:    # auxiliary variables: the un-haloed output that does not repeat them
                r = time_kernel_only(torch, runtime, wl, output="unknowns")
                others[wl]["output_unknowns_only"] = {"kernel_ms": r["ms"], "value": r["cell_updates_per_s"],
                                                      "achieved_GBs": r["algorithmic_GBs"], "frac": r["algorithmic_GBs"] / peak_gbs}

    # --- the same step under sustained load (informative): 1.2 s of back-to-back launches drive the board into its power
    # limit; the last third is timed, with the SM clock sampled meanwhile
    def sustained_leg(step_fn, per_step_ms):
        n_sus = max(300, int(1.2 / (per_step_ms * 1e-3)))
        s2 = ClockSampler(local_rank)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for i in range(n_sus):
            if i == 2 * n_sus // 3:
                torch.cuda.synchronize()
                s2.start()
                a.record(stream)
            step_fn()
        b.record(stream)
        torch.cuda.synchronize()
        sus_ms = a.elapsed_time(b) / (n_sus - 2 * n_sus // 3)
        return {"ms_per_step": sus_ms, "value": cells_per_step / (sus_ms * 1e-3), "unit": UNIT,
                "achieved_GBs": upd.algorithmic_bytes_per_patch * batch / (sus_ms * 1e-3) / 1e9,
                "frac_of_burst_peak": upd.algorithmic_bytes_per_patch * batch / (sus_ms * 1e-3) / 1e9 / measured_hbm_peak()[0],
                "launches": n_sus, "timed": n_sus - 2 * n_sus // 3, "clocks": s2.stop(),
                "note": "after ~0.8 s of back-to-back launches (board at its power limit, SM clock lowered)"}

    sustained = None
    if not args.no_sustained and world == 1:
        sustained = sustained_leg(step, kernel_ms)

    # --- the same workload with the opt-in fast arithmetic (EXAHYPE_FLAG_FAST_ARITHMETIC), informative: burst like the
    # headline (W warm-up + K steps of its own device-resident loop) and sustained; accuracy against the headline's output
    fast_leg = None
    if world == 1 and args.arithmetic == "reference" and not args.no_fast_leg and not args.no_sustained:
        time.sleep(1.0)                                        # let the board leave the power-limited state
        updf = runtime.PatchUpdate(model, dim, P, h, nr, na, dtype=dtype, dissipation=args.dissipation, output=args.output,
                                   kernel=args.kernel, arithmetic="fast")
        loopf = TimeLoop(dtype, None, args.cfl_dx, 0.01)
        q_fast = torch.empty_like(q_out)
        stepf = lambda: updf.step_loop(loopf, q_in, q_fast, lam_patch)
        for _ in range(args.warmup):
            stepf()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record(stream)
        for _ in range(args.steps):
            stepf()
        b.record(stream)
        torch.cuda.synchronize()
        fms = a.elapsed_time(b) / args.steps
        step()                                                 # the headline arithmetic on the same input and dt
        torch.cuda.synchronize()
        rel = float(((q_fast - q_out).abs().max() / q_out.abs().max()).item())
        fgbs = upd.algorithmic_bytes_per_patch * batch / (fms * 1e-3) / 1e9
        fast_leg = {"flag": "EXAHYPE_FLAG_FAST_ARITHMETIC (contracted FMAs, branch-free 1/x and sqrt; 3-deep ring + two staging buffers)",
                    "ms_per_step": fms, "value": cells_per_step / (fms * 1e-3), "achieved_GBs": fgbs,
                    "frac": fgbs / measured_hbm_peak()[0], "max_abs_err_over_max_abs_q_vs_reference_arithmetic": rel,
                    "bound": 1e-12, "sustained": sustained_leg(stepf, fms)}
        loopf.close()
        del q_fast

    if rank == 0:
        peak, peak_src = measured_hbm_peak()
        alg_bytes = upd.algorithmic_bytes_per_patch * batch          # per launch (= per GPU)
        achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
        info = upd.launch_info(batch)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": dtype, "data": "synthetic",
            "config": {"workload": desc, "patches_per_gpu": batch, "global_patches": batch * world,
                       "output": args.output, "dissipation": args.dissipation, "layout": "AoS (reference)", "kernel_variant": args.kernel,
                       "arithmetic": args.arithmetic + (" (bitwise equal to the reference's)" if args.arithmetic == "reference" else " (within 1e-12)"),
                       "parallelism": f"patch-sharded x{world}" + (", " + exchange_description(device_dt, fused, reducer) if world > 1 else ""),
                       "time_step": ("device-derived: dt(k+1) = cfl_dx / global lambda_max(k), produced and consumed on the "
                                     "device (exahype_cuda_fv_step_time_loop), no host sync between steps") if device_dt
                                    else "host constant",
                       "cfl_dx": args.cfl_dx if device_dt else None,
                       "l2": "inputs larger than L2: %.2f GB read + %.2f GB written per step per GPU"
                             % (q_in.numel() * q_in.element_size() / 1e9, q_out.numel() * q_out.element_size() / 1e9),
                       "kernel": info},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": recorded_traffic(args.workload + ("_unknowns" if args.output == "unknowns" else "")), "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": kernel_ms,
                         "kernel_ms_source": ("mean of per-launch CUDA event pairs inside the timed region" if args.step_events else
                                              "timed region / launches (two CUDA events around the K steps; includes launch gaps)"),
                         "frac_of_8TBs_nominal": achieved / 8000.0},
            "hbm_gbs_aggregate": achieved * world,
            "gpu_launches": int(launches),
            "clocks": clocks,
            "lambda_max": lam_global,
        }
        if checks is not None:
            line.update(checks)
        if e2e is not None:
            line["e2e"] = e2e
        if others is not None:
            line["other_workloads"] = others
        if sustained is not None:
            line["sustained"] = sustained
        if fast_leg is not None:
            line["fast_arithmetic"] = fast_leg
        if variants is not None:
            line["variants"] = variants
        if not args.no_cpu and world == 1:      # the CPU baseline belongs to the N = 1 line (bench contract)
            cpu_value, cpu_desc = cpu_reference_rate(args.workload, args.cpu_seconds, min(batch, args.cpu_sample))
            line["cpu_baseline"] = {"value": cpu_value, "unit": UNIT, **cpu_desc}
            # the reference's OWN compiled kernel next to the port (it is hard-wired to one 4x4 2-D patch and serial, so it
            # cannot run the workload: a reported figure, on its own shape)
            import oracle as O
            ref = O.reference_compiled_rate(2.0)
            if ref is not None:
                line["cpu_baseline"]["reference_compiled"] = {"value": ref[0], "unit": UNIT, "cores": 1, "sample": ref[1]}
        print(json.dumps(line), flush=True)

    if loop is not None:
        loop.close()
    if reducer is not None:
        reducer.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def exchange_description(device_dt, fused, reducer):
    if device_dt:
        return ("global time step per step: split-phase all-reduce(max) over NVLink peer memory (published by the patch "
                "kernel's last warp, consumed by the next launch's warps)")
    if reducer.backend != "peer":
        return "allreduce-max of lambda per step (NCCL)"
    return "allreduce-max of lambda per step (" + ("one-shot NVLink peer-memory exchange in the patch kernel's epilogue"
                                                   if fused else "one-shot NVLink peer-memory kernel") + ")"


def verify_time_loop(torch, dist, upd, reducer, args, q_in, q_out, lam_patch, shard, world, rank):
    """Three steps of a fresh device-resident loop, checked against independent paths (outside the timed region):
    the global lambda_max the loop consumed == torch.distributed MAX over the ranks' per-patch maxima; the dt it derived ==
    cfl_dx / that maximum in host arithmetic of the same type; sampled patches of this rank's q_out / lambda_patch ==
    the CPU oracle run with that dt, bit for bit.  max is exact and every rank divides the same two numbers, so this is
    the N-GPU == 1-GPU statement of SURVEY.md section 8e on hardware."""
    import numpy as np
    import oracle as O        # the checker, never the thing measured
    from exahype_b200.dist import TimeLoop
    npdt = np.float64 if upd.dtype == "f64" else np.float32
    loop = TimeLoop(upd.dtype, reducer, args.cfl_dx, 0.01)
    for _ in range(3):
        upd.step_loop(loop, q_in, q_out, lam_patch)
    loop.flush()
    torch.cuda.synchronize()
    hist = loop.history(0, 4)
    loop.close()
    lam_ref = lam_patch.max().reshape(1).clone()
    if world > 1:
        dist.all_reduce(lam_ref, op=dist.ReduceOp.MAX)            # NCCL: the independent path
    lam_ref = npdt(lam_ref.item())
    dt_ref = npdt(args.cfl_dx) / lam_ref
    ok_lambda = bool(hist[1, 1] == lam_ref and hist[2, 1] == lam_ref and hist[3, 1] == lam_ref and hist[0, 1] == 0)
    ok_dt = bool(hist[0, 0] == npdt(0.01) and hist[1, 0] == dt_ref and hist[2, 0] == dt_ref and hist[3, 0] == dt_ref)
    # sampled patches of this shard against the oracle with the dt the device derived
    n = q_in.shape[0]
    picks = sorted(set(list(range(min(8, n))) + list(range(n // 2, min(n, n // 2 + 8))) + list(range(max(0, n - 8), n))))
    cfg = O.OracleConfig(dim=upd.dim, patch_size=upd.patch_size, halo=upd.halo_size, n_real=upd.n_real, n_aux=upd.n_aux,
                         model={"euler": O.MODEL_EULER, "swe": O.MODEL_SWE, "swe_source": O.MODEL_SWE_SOURCE}[upd.model],
                         diss=O.DISS_ALL if upd.dissipation == "all" else O.DISS_VAR0)
    idx = torch.tensor(picks, device=q_in.device)
    want = q_in[idx].cpu().numpy()
    # the shard is a slice of ONE global batch: the same bits the oracle generates for those global patch numbers
    ok_input = all(np.array_equal(want[i], O.fill_synthetic(cfg, 1, first_patch=shard.first + p, dtype=npdt)[0])
                   for i, p in enumerate(picks[:4]))
    lam_o, _ = O.step(cfg, want, float(dt_ref), nthreads=2)
    got = q_out[idx].cpu().numpy()
    if upd.output in ("unhaloed", "unknowns"):
        h = upd.halo_size
        sl = (slice(None),) + (slice(h, -h),) * upd.dim + (slice(0, upd.n_real if upd.output == "unknowns" else None),)
        want = want[sl]
    if upd.arithmetic == "fast":       # opt-in arithmetic: the north star's 1e-12 bound instead of bit equality
        tol = 1e-12 if upd.dtype == "f64" else 2e-6
        ok_q = bool(np.abs(got - want).max() <= tol * np.abs(want).max()) and \
            bool(np.allclose(lam_patch[idx].cpu().numpy(), lam_o, rtol=tol, atol=0))
    else:
        ok_q = bool(np.array_equal(got, want)) and bool(np.array_equal(lam_patch[idx].cpu().numpy(), lam_o))
    flags = torch.tensor([ok_lambda, ok_dt, ok_input, ok_q], dtype=torch.int32, device=q_in.device)
    if world > 1:
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    ok_lambda, ok_dt, ok_input, ok_q = (bool(x) for x in flags.tolist())
    timed_out = bool(reducer.timed_out()) if reducer is not None else False
    return {"multi_gpu_bitwise": bool(ok_lambda and ok_dt and ok_input and ok_q and not timed_out),
            "time_loop_check": {"global_lambda_equals_nccl_max": ok_lambda, "dt_equals_host_division": ok_dt,
                                "shard_input_equals_global_batch_slice": ok_input,
                                "sampled_patches_equal_oracle_bitwise": ok_q, "patches_sampled_per_rank": len(picks),
                                "dt_device": float(hist[1, 0]), "exchange_timed_out": timed_out}}


def synthetic_on_device(torch, upd, first_patch: int, n_patches: int, tdt, device="cuda"):
    """SplitMix64-based admissible state of SURVEY.md 8d, generated by the library's own kernel
    (exahype_cuda_fill_synthetic) so a 1.3 GB shard does not cross PCIe.  Bit-identical to oracle.fill_synthetic
    (tests/test_gpu_parity.py::test_device_generator_equals_oracle_fill)."""
    q = torch.empty(upd.in_shape(n_patches), dtype=tdt, device=device)
    return upd.fill_synthetic(q, first_patch)


def time_kernel_only(torch, runtime, wl, output="unhaloed", diss="var0", kern="auto", steps=10):
    """Kernel-only timing of one committed variant on its BASELINE batch (same timing hygiene as the main loop: warm-up,
    CUDA events on the launching stream, inputs larger than L2)."""
    model, dim, P, h, nr, na, dtype, batch, _ = WORKLOADS[wl]
    tdt = torch.float64 if dtype == "f64" else torch.float32
    upd = runtime.PatchUpdate(model, dim, P, h, nr, na, dtype=dtype, dissipation=diss, output=output, kernel=kern)
    q_in = synthetic_on_device(torch, upd, 0, batch, tdt)
    q_out = torch.empty(upd.out_shape(batch), dtype=tdt, device="cuda")
    lam = torch.zeros(1, dtype=tdt, device="cuda")
    for _ in range(3):
        upd.step(q_in, q_out, 0.01, None, lam)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(steps):
        upd.step(q_in, q_out, 0.01, None, lam)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    gbs = upd.algorithmic_bytes_per_patch * batch / (ms * 1e-3) / 1e9
    return {"ms": ms, "cell_updates_per_s": batch * P ** dim / (ms * 1e-3), "algorithmic_GBs": gbs}


def time_in_place(torch, runtime, wl, diss="var0", steps=10):
    """The reference's own call shape -- ``time_step(Q, dt)``: haloed, IN PLACE (q_out == q_in) -- kernel only.  The state is
    restored from a pristine copy before every launch, outside that launch's events (repeated in-place steps would leave
    the admissible synthetic input)."""
    model, dim, P, h, nr, na, dtype, batch, _ = WORKLOADS[wl]
    tdt = torch.float64 if dtype == "f64" else torch.float32
    upd = runtime.PatchUpdate(model, dim, P, h, nr, na, dtype=dtype, dissipation=diss, output="haloed")
    q0 = synthetic_on_device(torch, upd, 0, batch, tdt)
    q = q0.clone()
    lam = torch.zeros(1, dtype=tdt, device="cuda")
    ms = 0.0
    for it in range(3 + steps):
        q.copy_(q0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        upd.step(q, None, 0.01, None, lam)
        b.record()
        torch.cuda.synchronize()
        if it >= 3:
            ms += a.elapsed_time(b) / steps
    gbs = upd.algorithmic_bytes_per_patch * batch / (ms * 1e-3) / 1e9
    return {"ms": ms, "cell_updates_per_s": batch * P ** dim / (ms * 1e-3), "algorithmic_GBs": gbs}


def time_variants(torch, runtime, args):
    """Kernel-only timings of the other committed variants / workloads (informative; same timing hygiene)."""
    res = {}
    for wl in ("c3", "c2", "c4", "c4f32", "c1"):
        dim = WORKLOADS[wl][1]
        outputs = ("unhaloed", "haloed") + (("unknowns",) if WORKLOADS[wl][5] > 0 else ())
        for output, diss, kern in [(o, d, k) for o in outputs for d in ("var0", "all")
                                   for k in (("auto", "cell") if dim == 3 else ("auto",))]:
            res[f"{wl}/{output}/{diss}" + ("/cell-kernel" if kern == "cell" else "")] = \
                time_kernel_only(torch, runtime, wl, output, diss, kern)
        if wl != "c1":
            for diss in ("var0", "all"):
                res[f"{wl}/haloed-in-place/{diss} (the reference's call shape)"] = time_in_place(torch, runtime, wl, diss)
    # the CellData form (per-patch pointers + per-patch dt) on the headline workload: same kernel, gathered addressing
    for wl in ("c3", "c2"):
        model, dim, P, h, nr, na, dtype, batch, _ = WORKLOADS[wl]
        tdt = torch.float64
        upd = runtime.PatchUpdate(model, dim, P, h, nr, na, dtype=dtype, output="unhaloed")
        q_in = synthetic_on_device(torch, upd, 0, batch, tdt)
        q_out = torch.empty(upd.out_shape(batch), dtype=tdt, device="cuda")
        perm = torch.randperm(batch, device="cuda")
        per_in, per_out = q_in[0].numel() * 8, q_out[0].numel() * 8
        in_ptrs = q_in.data_ptr() + perm * per_in
        out_ptrs = q_out.data_ptr() + perm * per_out
        dts = torch.full((batch,), 0.01, dtype=tdt, device="cuda")
        lam = torch.zeros(1, dtype=tdt, device="cuda")
        for _ in range(3):
            upd.step_cell_data(in_ptrs, out_ptrs, dt_patch=dts, lambda_max=lam)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(10):
            upd.step_cell_data(in_ptrs, out_ptrs, dt_patch=dts, lambda_max=lam)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 10
        res[f"{wl}/unhaloed/var0/cell-data (permuted patch pointers, per-patch dt)"] = {
            "ms": ms, "cell_updates_per_s": batch * P ** dim / (ms * 1e-3),
            "algorithmic_GBs": upd.algorithmic_bytes_per_patch * batch / (ms * 1e-3) / 1e9}
        del q_in, q_out
    return res


if __name__ == "__main__":
    main()
