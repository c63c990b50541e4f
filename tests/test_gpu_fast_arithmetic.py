"""EXAHYPE_FLAG_FAST_ARITHMETIC (PatchUpdate(arithmetic='fast')): contracted multiply-adds and a branch-free reciprocal /
square root.  Not bitwise -- the contract is BASELINE.json's bound, 1e-12 relative in fp64 (max norm over the batch,
relative to the largest state value), asserted here against the oracle for every committed fast instantiation, both
dissipation variants, both output forms, and for the device-resident time loop.  fp32: the fp32 bound of
tests/test_gpu_parity.py (2e-6 of the largest state value)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL_F64 = 1e-12
RTOL_F32 = 2e-6

SHAPES = [("euler", 3, 8, 5, 0, "f64", 700), ("euler", 2, 16, 4, 0, "f64", 300), ("swe", 2, 32, 3, 1, "f64", 120),
          ("swe", 2, 32, 3, 1, "f32", 120)]


@pytest.fixture(scope="module")
def torch():
    import torch
    assert torch.cuda.is_available()
    return torch


@pytest.fixture(scope="module")
def rt():
    from exahype_b200 import runtime
    return runtime


def cfg_of(oracle, upd):
    return oracle.OracleConfig(dim=upd.dim, patch_size=upd.patch_size, halo=upd.halo_size, n_real=upd.n_real,
                               n_aux=upd.n_aux, model={"euler": oracle.MODEL_EULER, "swe": oracle.MODEL_SWE, "swe_source": oracle.MODEL_SWE_SOURCE}[upd.model],
                               diss=oracle.DISS_ALL if upd.dissipation == "all" else oracle.DISS_VAR0)


@pytest.mark.parametrize("model,dim,P,nr,na,dtype,B", SHAPES)
@pytest.mark.parametrize("diss", ["var0", "all"])
@pytest.mark.parametrize("output", ["haloed", "unhaloed"])
def test_fast_arithmetic_within_the_stated_bound(torch, rt, oracle, model, dim, P, nr, na, dtype, B, diss, output):
    upd = rt.PatchUpdate(model, dim, P, 1, nr, na, dtype=dtype, dissipation=diss, output=output, arithmetic="fast")
    npdt = np.float64 if dtype == "f64" else np.float32
    tol = RTOL_F64 if dtype == "f64" else RTOL_F32
    q0 = oracle.fill_synthetic(cfg_of(oracle, upd), B, dtype=npdt)
    want = q0.copy()
    lam_o, lmax_o = oracle.step(cfg_of(oracle, upd), want, 0.01, nthreads=4)
    q = torch.from_numpy(q0).cuda()
    out = q if output == "haloed" else torch.full(upd.out_shape(B), 9.0, dtype=q.dtype, device="cuda")
    lam = torch.zeros(B, dtype=q.dtype, device="cuda")
    lmax = torch.zeros(1, dtype=q.dtype, device="cuda")
    upd.step(q, out, 0.01, lam, lmax)
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    if output == "unhaloed":
        want = want[(slice(None),) + (slice(1, -1),) * dim + (slice(None),)]
    err = np.abs(got.astype(np.float64) - want.astype(np.float64)).max() / np.abs(want).max()
    assert err <= tol, f"fast arithmetic off by {err:.3e} (bound {tol:g})"
    np.testing.assert_allclose(lam.cpu().numpy(), lam_o, rtol=tol, atol=0)
    assert abs(float(lmax.item()) - float(lmax_o)) <= tol * float(lmax_o)
    if output == "haloed":                   # halo cells and aux variables are copied through untouched, bit for bit
        halo = np.ones(q0.shape[1:-1], dtype=bool)
        halo[(slice(1, -1),) * dim] = False
        assert np.array_equal(got[:, halo, :], q0[:, halo, :])
        if na:
            assert np.array_equal(got[..., nr:], q0[..., nr:])


def test_fast_arithmetic_is_a_different_kernel_and_falls_back_where_none_exists(torch, rt, oracle):
    """The flag selects the ArithFast instantiation where one is committed (results differ from the reference's in the last
    bits); a shape without one keeps the reference arithmetic, bit for bit."""
    for shape, expect_equal in ((("euler", 3, 8, 1, 5, 0), False), (("euler", 3, 4, 1, 5, 0), True)):
        ref = rt.PatchUpdate(*shape, output="haloed")
        fast = rt.PatchUpdate(*shape, output="haloed", arithmetic="fast")
        q0 = oracle.fill_synthetic(cfg_of(oracle, ref), 64)
        a, b = torch.from_numpy(q0).cuda(), torch.from_numpy(q0).cuda()
        ref.step(a, a, 0.01)
        fast.step(b, b, 0.01)
        torch.cuda.synchronize()
        assert bool(torch.equal(a, b)) == expect_equal
        if not expect_equal:
            assert ref.launch_info(64)["smem_bytes"] != fast.launch_info(64)["smem_bytes"]   # 4-deep/1 vs 3-deep/2 staging


def test_fast_arithmetic_time_loop(torch, rt, oracle):
    """The device-resident loop with the fast kernels: dt of the later steps derives from lambda_max of the fast
    arithmetic (within the bound of the oracle's), and so does the state.  The input stays fixed (q_in -> q_out, as in
    bench.py): evolving this random state in place is ill-conditioned, which would test the data, not the kernel."""
    from exahype_b200.dist import TimeLoop
    upd = rt.PatchUpdate("euler", 3, 8, 1, 5, 0, output="haloed", arithmetic="fast")
    cfg = cfg_of(oracle, upd)
    q0 = oracle.fill_synthetic(cfg, 200)
    cfl_dx = 0.05
    probe = q0.copy()
    _, lmax = oracle.step(cfg, probe, 0.01, nthreads=4)
    dt = cfl_dx / lmax                                  # what every step after the first uses
    want = q0.copy()
    oracle.step(cfg, want, float(dt), nthreads=4)
    q = torch.from_numpy(q0).cuda()
    out = q.clone()
    loop = TimeLoop("f64", None, cfl_dx, 0.01)
    for _ in range(3):
        upd.step_loop(loop, q, out)
    loop.flush()
    torch.cuda.synchronize()
    h = loop.history(0, 4)
    assert h[0, 0] == 0.01 and abs(h[2, 0] - dt) <= RTOL_F64 * dt and abs(h[3, 0] - dt) <= RTOL_F64 * dt
    assert np.abs(out.cpu().numpy() - want).max() <= 2 * RTOL_F64 * np.abs(want).max()
    loop.close()
