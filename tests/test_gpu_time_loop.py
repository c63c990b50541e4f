"""The global admissible time step on the device and the peer-memory exchange under it.

One GPU is enough: `LocalPeerGroup` puts several ranks on the current device (plain device pointers instead of CUDA IPC
mappings, each rank on its own stream), so the mailbox protocol -- sequence parity, ticket reset, blocking and
split-phase forms, the empty-shard path, the timeout path -- runs exactly the kernels a multi-GPU job runs.  The
multi-GPU job itself is checked by bench.py at N > 1 (`multi_gpu_bitwise`).

Contract: every step of an R-rank loop uses bit for bit the dt of the 1-rank loop over the whole batch, and both equal
the oracle driven by the same rule  dt_{k+1} = cfl_dx / lambda_max(step k)  on the host.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CFL_DX = 0.4 * 0.125
DT0 = 0.01


@pytest.fixture(scope="module")
def torch():
    import torch
    assert torch.cuda.is_available()
    return torch


@pytest.fixture(scope="module")
def rt():
    from exahype_b200 import runtime
    return runtime


def oracle_cfg(oracle, upd):
    return oracle.OracleConfig(dim=upd.dim, patch_size=upd.patch_size, halo=upd.halo_size, n_real=upd.n_real,
                               n_aux=upd.n_aux, model={"euler": oracle.MODEL_EULER, "swe": oracle.MODEL_SWE, "swe_source": oracle.MODEL_SWE_SOURCE}[upd.model],
                               diss=oracle.DISS_ALL if upd.dissipation == "all" else oracle.DISS_VAR0)


def oracle_loop(oracle, upd, q0, steps, npdt):
    """The same time loop on the host: returns the states after each step, dt used per step, lambda_max per step."""
    cfg = oracle_cfg(oracle, upd)
    q = q0.copy()
    dt = npdt(DT0)
    states, dts, lams = [], [], []
    for _ in range(steps):
        _, lmax = oracle.step(cfg, q, float(dt), nthreads=4)
        states.append(q.copy()); dts.append(dt); lams.append(npdt(lmax))
        if lams[-1] > 0:
            dt = npdt(CFL_DX) / lams[-1]
    return states, dts, lams


SHAPES = [
    # model, dim, P, nr, na, dtype, dissipation, patches      kernel family
    ("euler", 3, 8, 5, 0, "f64", "var0", 150),              # warp per patch: exchange inside the patch kernel
    ("euler", 3, 8, 5, 0, "f64", "all", 37),
    ("euler", 3, 8, 5, 0, "f32", "var0", 64),
    ("euler", 2, 16, 4, 0, "f64", "var0", 100),             # row marching: consume in the kernel, one-warp publish behind it
    ("swe", 2, 32, 3, 1, "f32", "all", 40),
    ("euler", 3, 4, 5, 0, "f64", "var0", 50),               # warp groups
    ("euler", 2, 3, 4, 0, "f64", "var0", 1000),             # thread per cell
]


@pytest.mark.parametrize("model,dim,P,nr,na,dtype,diss,B", SHAPES)
def test_one_rank_loop_equals_host_loop(torch, rt, oracle, model, dim, P, nr, na, dtype, diss, B):
    from exahype_b200.dist import TimeLoop
    upd = rt.PatchUpdate(model, dim, P, 1, nr, na, dtype=dtype, dissipation=diss, output="haloed")
    npdt = np.float64 if dtype == "f64" else np.float32
    q0 = oracle.fill_synthetic(oracle_cfg(oracle, upd), B, dtype=npdt)
    steps = 4
    states, dts, lams = oracle_loop(oracle, upd, q0, steps, npdt)
    q = torch.from_numpy(q0).cuda()
    lam_patch = torch.zeros(B, dtype=q.dtype, device="cuda")
    loop = TimeLoop(dtype, None, CFL_DX, DT0)
    launches = rt.launch_count()
    for k in range(steps):
        upd.step_loop(loop, q, q, lam_patch)          # in place, like the reference's time_step(Q, dt)
    fused = rt.launch_count() - launches == steps       # one launch per step where the kernel runs the exchange itself
    assert fused == (dim == 3 and P == 8)
    loop.flush()
    torch.cuda.synchronize()
    assert np.array_equal(q.cpu().numpy(), states[-1]), "state after the device-resident loop differs from the host loop"
    h = loop.history(0, steps + 1)
    assert np.array_equal(h[:steps, 0], np.array(dts, dtype=npdt)), "dt sequence"
    assert np.array_equal(h[1:steps + 1, 1], np.array(lams, dtype=npdt)), "global lambda_max sequence"
    assert np.array_equal(h[:steps, 2], np.array(lams, dtype=npdt)), "device lambda_max sequence"
    assert h[0, 1] == 0 and h[steps, 0] == npdt(CFL_DX) / lams[-1]
    assert float(lam_patch.max().item()) == float(lams[-1])
    # after the flush the loop continues from the device scalar
    upd.step_loop(loop, q, q, lam_patch)
    torch.cuda.synchronize()
    assert loop.history(steps, 1)[0, 0] == h[steps, 0]
    loop.close()


@pytest.mark.parametrize("world,shape", [(2, 0), (3, 0), (2, 3), (4, 1)])
def test_ranks_on_one_device_equal_one_rank(torch, rt, oracle, world, shape):
    """R ranks, each with a contiguous shard on its own stream, against the host loop over the whole batch: identical dt
    every step, identical state.  One of the shards is empty in the (4, ...) case."""
    from exahype_b200.dist import LocalPeerGroup, PatchSharding, TimeLoop
    model, dim, P, nr, na, dtype, diss, _ = SHAPES[shape]
    per_rank = 48
    B = per_rank * world if world != 4 else 3          # 4 ranks over 3 patches: rank shards of 0 / 1 patches
    upd = rt.PatchUpdate(model, dim, P, 1, nr, na, dtype=dtype, dissipation=diss, output="haloed")
    npdt = np.float64 if dtype == "f64" else np.float32
    q0 = oracle.fill_synthetic(oracle_cfg(oracle, upd), B, dtype=npdt)
    steps = 6
    states, dts, lams = oracle_loop(oracle, upd, q0, steps, npdt)

    group = LocalPeerGroup(world)
    shards = [PatchSharding(B, world, r) for r in range(world)]
    qs = [torch.from_numpy(q0[s.slice()].copy()).cuda() for s in shards]
    streams = [torch.cuda.Stream() for _ in range(world)]
    loops = [TimeLoop(dtype, group[r], CFL_DX, DT0) for r in range(world)]
    torch.cuda.synchronize()
    for k in range(steps):
        for r in range(world):
            with torch.cuda.stream(streams[r]):
                upd.step_loop(loops[r], qs[r], qs[r])
    for r in range(world):
        with torch.cuda.stream(streams[r]):
            loops[r].flush()
    torch.cuda.synchronize()
    assert not any(group[r].timed_out() for r in range(world))
    for r, s in enumerate(shards):
        assert np.array_equal(qs[r].cpu().numpy(), states[-1][s.slice()]), f"rank {r}: shard differs from the 1-rank run"
        h = loops[r].history(0, steps + 1)
        assert np.array_equal(h[:steps, 0], np.array(dts, dtype=npdt)), f"rank {r}: dt sequence"
        assert np.array_equal(h[1:, 1], np.array(lams, dtype=npdt)), f"rank {r}: global lambda_max"
    for l in loops:
        l.close()
    group.close()


@pytest.mark.parametrize("world", [2, 5])
def test_blocking_exchange_many_rounds(torch, rt, world):
    """Stand-alone one-shot all-reduce(max): 1200 exchanges back to back, a different winner every round (sequence
    parity of the two mailbox slots, no lost or stale value)."""
    from exahype_b200.dist import LocalPeerGroup
    group = LocalPeerGroup(world)
    rounds = 1200
    rng = np.random.default_rng(7)
    vals = rng.random((rounds, world)) + 0.5
    dev = [torch.from_numpy(vals[:, r].copy()).cuda() for r in range(world)]
    streams = [torch.cuda.Stream() for _ in range(world)]
    torch.cuda.synchronize()
    for k in range(rounds):
        for r in range(world):
            with torch.cuda.stream(streams[r]):
                group[r].allreduce_max(dev[r][k:k + 1])
    torch.cuda.synchronize()
    want = vals.max(axis=1)
    for r in range(world):
        assert np.array_equal(dev[r].cpu().numpy(), want)
        assert not group[r].timed_out()
    group.close()


def test_blocking_exchange_in_the_patch_kernel(torch, rt, oracle):
    """exahype_cuda_fv_step_allreduce on two ranks of one device: fused epilogue (ticket reset across launches), the
    empty-shard fallback on one rank, results equal to the oracle's global maximum."""
    from exahype_b200.dist import LocalPeerGroup
    upd = rt.PatchUpdate("euler", 3, 8, 1, 5, 0, output="haloed")
    cfg = oracle_cfg(oracle, upd)
    B = 40
    q0 = oracle.fill_synthetic(cfg, B)
    lam_o, lmax_o = oracle.step(cfg, q0.copy(), 0.01, nthreads=4)
    for split in (B // 2, B):                       # second case: rank 1 owns nothing
        group = LocalPeerGroup(2)
        parts = [q0[:split], q0[split:]]
        streams = [torch.cuda.Stream() for _ in range(2)]
        lam = [torch.zeros(1, dtype=torch.float64, device="cuda") for _ in range(2)]
        for rep in range(5):
            qs = [torch.from_numpy(p.copy()).cuda() for p in parts]
            torch.cuda.synchronize()
            for r in range(2):
                with torch.cuda.stream(streams[r]):
                    upd.step(qs[r], qs[r], 0.01, None, lam[r], reducer=group[r])
            torch.cuda.synchronize()
            assert float(lam[0].item()) == float(lam[1].item()) == float(lmax_o)
        assert not group[0].timed_out() and not group[1].timed_out()
        group.close()


def test_timeout_poisons_and_sticks(torch, rt):
    """A peer that never arrives: the waiting rank gets NaN (never a silently rank-local value), the flag is visible
    without synchronising, and every later call on the reducer fails with TIMEOUT."""
    from exahype_b200.dist import LocalPeerGroup, TimeLoop
    group = LocalPeerGroup(2)
    group[0].set_timeout(0.02)
    v = torch.full((1,), 3.0, dtype=torch.float64, device="cuda")
    group[0].allreduce_max(v)                        # rank 1 never calls
    torch.cuda.synchronize()
    assert np.isnan(v.item())
    assert group[0].timed_out()
    with pytest.raises(rt.ExaHyPECudaError) as e:
        group[0].allreduce_max(v)
    assert e.value.code == -6
    group.close()

    # the same in the time loop: rank 0 steps twice alone -> its second step consumes an exchange rank 1 never joined
    group = LocalPeerGroup(2)
    group[0].set_timeout(0.02)
    upd = rt.PatchUpdate("euler", 3, 8, 1, 5, 0, output="haloed")
    q = upd.fill_synthetic(torch.empty(upd.in_shape(16), dtype=torch.float64, device="cuda"), 0)
    loop = TimeLoop("f64", group[0], CFL_DX, DT0)
    upd.step_loop(loop, q, q)
    upd.step_loop(loop, q, q)
    torch.cuda.synchronize()
    assert group[0].timed_out()
    assert bool(torch.isnan(q).any()), "a timed-out exchange must poison the step, not fall back to a local dt"
    with pytest.raises(rt.ExaHyPECudaError) as e:
        upd.step_loop(loop, q, q)
    assert e.value.code == -6
    loop.close()
    group.close()


def test_reducer_refuses_blocking_calls_while_a_loop_exchange_is_pending(torch, rt):
    from exahype_b200.dist import LocalPeerGroup, TimeLoop
    group = LocalPeerGroup(1)
    upd = rt.PatchUpdate("euler", 3, 8, 1, 5, 0, output="haloed")
    q = upd.fill_synthetic(torch.empty(upd.in_shape(8), dtype=torch.float64, device="cuda"), 0)
    loop = TimeLoop("f64", group[0], CFL_DX, DT0)
    upd.step_loop(loop, q, q)
    v = torch.ones(1, dtype=torch.float64, device="cuda")
    with pytest.raises(rt.ExaHyPECudaError):
        group[0].allreduce_max(v)
    loop.flush()
    group[0].allreduce_max(v)
    torch.cuda.synchronize()
    assert v.item() == 1.0
    loop.close()
    group.close()


def test_exchange_trace_is_monotonic(torch, rt):
    """The globaltimer stamps bench.py / scripts use to attribute the exchange cost: begin <= wait <= ... per step."""
    from exahype_b200.dist import LocalPeerGroup, TimeLoop
    group = LocalPeerGroup(1)
    group[0].enable_trace(64)
    upd = rt.PatchUpdate("euler", 3, 8, 1, 5, 0, output="unhaloed")
    q = upd.fill_synthetic(torch.empty(upd.in_shape(2048), dtype=torch.float64, device="cuda"), 0)
    out = torch.empty(upd.out_shape(2048), dtype=torch.float64, device="cuda")
    loop = TimeLoop("f64", group[0], CFL_DX, DT0)
    for _ in range(10):
        upd.step_loop(loop, q, out)
    tr = group[0].read_trace(1, 10).astype(np.int64)
    # exchange s: published by step s-1 (LAST_WARP, PUBLISHED), consumed by step s (WAIT_BEGIN, WAIT_END)
    for s in range(9):
        last_warp, published, wait_b, wait_e = tr[s, 3], tr[s, 4], tr[s, 1], tr[s, 2]
        assert 0 < last_warp <= published
        if s < 9 and wait_b:
            assert published <= wait_e and wait_b <= wait_e
    loop.close()
    group.close()
