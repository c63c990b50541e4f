"""Parity of the CUDA path (through the C ABI) against the CPU oracle -- bit-exact in fp64 and fp32.

The oracle reproduces the reference's compiled kernel bit for bit (tests/test_oracle_golden.py), so equality with
the oracle is equality with the reference's arithmetic.  The stated contract (BASELINE.json north_star) is 1e-12
relative in fp64; these tests hold the stronger property, and `TOL` below is only used where a looser
statement is the point (fp32 against fp64).
"""
import dataclasses
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL_FP64 = 1e-12          # the north-star bound; the tests assert bitwise equality, which implies it
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def torch():
    import torch
    assert torch.cuda.is_available()
    return torch


@pytest.fixture(scope="module")
def rt():
    from exahype_b200 import runtime
    assert runtime.device_count() >= 1
    return runtime


GUARD = 512          # elements of sentinel on either side of every device buffer the kernels write
SENTINEL = -977.0


def guarded(torch, shape, tdt, fill=None):
    """A device tensor of `shape` carved out of a larger allocation whose margins hold a sentinel: a kernel that writes
    outside its buffers (compute-sanitizer is closed on this GPU pool) trips check_guards()."""
    n = int(np.prod(shape))
    raw = torch.full((n + 2 * GUARD,), SENTINEL, dtype=tdt, device="cuda")
    view = raw[GUARD:GUARD + n].view(shape)
    if fill is not None:
        view.fill_(fill)
    return raw, view


def check_guards(raws):
    for raw in raws:
        assert bool((raw[:GUARD] == SENTINEL).all()) and bool((raw[-GUARD:] == SENTINEL).all()), "write outside a buffer"


def gpu_step(torch, upd, q_np, dt, out=None, want_lambda=True):
    """Runs one step on the device; returns (q_out numpy, lambda_patch numpy, lambda_max).  Every buffer sits between
    sentinel guard bands that are verified after the launch."""
    tdt = torch.float64 if upd.dtype == "f64" else torch.float32
    n = q_np.shape[0]
    raw_q, q = guarded(torch, q_np.shape, tdt)
    q.copy_(torch.from_numpy(q_np))
    raw_l, lam = guarded(torch, (n,), tdt, -1.0)
    raw_m, lmax = guarded(torch, (1,), tdt, -1.0)
    raws = [raw_q, raw_l, raw_m]
    if upd.output in ("unhaloed", "unknowns"):
        raw_o, out = guarded(torch, upd.out_shape(n), tdt, 7.0)
        raws.append(raw_o)
    res = upd.step(q, out, dt, lam if want_lambda else None, lmax if want_lambda else None)
    torch.cuda.synchronize()
    check_guards(raws)
    return res.cpu().numpy(), lam.cpu().numpy(), float(lmax.item())


def oracle_cfg(oracle, upd):
    return oracle.OracleConfig(dim=upd.dim, patch_size=upd.patch_size, halo=upd.halo_size, n_real=upd.n_real,
                               n_aux=upd.n_aux, model={"euler": oracle.MODEL_EULER, "swe": oracle.MODEL_SWE, "swe_source": oracle.MODEL_SWE_SOURCE}[upd.model],
                               diss=oracle.DISS_ALL if upd.dissipation == "all" else oracle.DISS_VAR0)


def interior(upd, a):
    h = upd.halo_size
    sl = (slice(None),) + (slice(h, -h),) * upd.dim
    return a[sl]


def assert_bitwise(a, b, what=""):
    a = np.ascontiguousarray(a); b = np.ascontiguousarray(b)
    same = a.view(np.uint8) == b.view(np.uint8)
    if not same.all():
        bad = np.argwhere(a != b)
        rel = np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)
        raise AssertionError(f"{what}: {len(bad)} values differ, first at {bad[0]}, max rel diff {rel:.3e}")


# ----------------------------------------------------------------------------------------- goldens
@pytest.mark.parametrize("diss", ["var0", "all"])
@pytest.mark.parametrize("dt", [1.0, 0.01])
def test_config_c1_sin_input_golden_g1(torch, rt, oracle, diss, dt):
    """BASELINE config C1: 2-D Euler 3x3+1, 1000 patches (ragged last tile: 1000 = 35*28 + 20)."""
    upd = rt.PatchUpdate("euler", 2, 3, 1, 4, 0, dissipation=diss)
    cfg = oracle_cfg(oracle, upd)
    q0 = oracle.fill_sin(cfg, 1000)
    want = q0.copy()
    lam_o, lmax_o = oracle.step(cfg, want, dt)
    got, lam, lmax = gpu_step(torch, upd, q0, dt)
    assert_bitwise(got, want, "C1")
    assert_bitwise(lam, lam_o, "lambda_patch")
    assert lmax == lmax_o == 1.1326339818690889
    expected = {("var0", 1.0): "1334ede917c43f96", ("all", 1.0): "aa5c6cb6b3c1c18e",
                ("var0", 0.01): "fbacf0e93e60875b", ("all", 0.01): "369e22f97aff515c"}[(diss, dt)]
    assert oracle.fnv1a64(got) == expected


def test_reference_kernel_shape_g0(torch, rt, oracle):
    """Shape of the reference's committed kernel (2-D, 4x4+1, 5+5 variables).  The committed file's transposed
    ranges read uninitialised rows; on the cells those rows cannot reach, the CUDA result equals the output of the
    reference's own compiled kernel bit for bit (fixture written by tests/golden/make_golden.py)."""
    upd = rt.PatchUpdate("euler", 2, 4, 1, 5, 5)
    cfg = oracle_cfg(oracle, upd)
    with open(os.path.join(HERE, "golden", "g0_reference_kernel.json")) as f:
        g0 = json.load(f)
    bits = lambda hx: np.array([int(h, 16) for h in hx], dtype=np.uint64).view(np.float64)
    q0 = bits(g0["input"]).reshape(1, 6, 6, 10).copy()
    ref_out = bits(g0["output"]).reshape(1, 6, 6, 10)
    got, lam, lmax = gpu_step(torch, upd, q0, g0["dt"])
    assert_bitwise(got[0, 2:4, 2:4], ref_out[0, 2:4, 2:4], "inner cells vs compiled reference")
    assert_bitwise(got[..., 5:], q0[..., 5:], "aux pass through")
    want = q0.copy(); oracle.step(cfg, want, g0["dt"])
    assert_bitwise(got, want, "G0h")
    assert oracle.fnv1a64(got) == "5c1d83d32c1d0f28"


@pytest.mark.parametrize("diss,fill,expected", [("var0", "sin", "53201509629d2991"), ("all", "sin", "78219e9b15eba6f7"),
                                                ("var0", "syn", "f8486ef10dfe78d6"), ("all", "syn", "c72ba3f4ebf895ec")])
def test_euler3d_goldens_g3(torch, rt, oracle, diss, fill, expected):
    upd = rt.PatchUpdate("euler", 3, 8, 1, 5, 0, dissipation=diss)
    cfg = oracle_cfg(oracle, upd)
    q0 = oracle.fill_sin(cfg, 4) if fill == "sin" else oracle.fill_synthetic(cfg, 4)
    want = q0.copy(); lam_o, lmax_o = oracle.step(cfg, want, 0.01)
    got, lam, lmax = gpu_step(torch, upd, q0, 0.01)
    assert_bitwise(got, want, "G3")
    assert oracle.fnv1a64(got) == expected
    assert_bitwise(lam, lam_o, "lambda")
    assert lmax == lmax_o
    np.testing.assert_allclose(got, want, rtol=TOL_FP64, atol=0)


@pytest.mark.parametrize("diss,expected", [("var0", "1856adbc23fec8cd"), ("all", "8e53779668580e14")])
def test_euler2d_golden_g2r(torch, rt, oracle, diss, expected):
    upd = rt.PatchUpdate("euler", 2, 16, 1, 4, 0, dissipation=diss)
    cfg = oracle_cfg(oracle, upd)
    q0 = oracle.fill_synthetic(cfg, 8)
    got, lam, lmax = gpu_step(torch, upd, q0, 0.01)
    assert oracle.fnv1a64(got) == expected
    assert lmax == 2.1332213875447144


@pytest.mark.parametrize("model,dim,P,nr,na,dtype", [("euler", 3, 8, 5, 0, "f64"), ("euler", 2, 16, 4, 0, "f64"),
                                                      ("swe", 2, 32, 3, 1, "f64"), ("swe", 2, 32, 3, 1, "f32"),
                                                      ("euler", 2, 3, 4, 0, "f64"), ("euler", 2, 4, 5, 5, "f64")])
def test_device_generator_equals_oracle_fill(torch, rt, oracle, model, dim, P, nr, na, dtype):
    """bench.py generates its input with exahype_cuda_fill_synthetic: it must be the oracle's synthetic state bit for
    bit, shard by shard (so the benchmark runs on exactly the data the parity tests cover)."""
    upd = rt.PatchUpdate(model, dim, P, 1, nr, na, dtype=dtype)
    cfg = oracle_cfg(oracle, upd)
    tdt = torch.float64 if dtype == "f64" else torch.float32
    npdt = np.float64 if dtype == "f64" else np.float32
    q = torch.zeros(upd.in_shape(9), dtype=tdt, device="cuda")
    got = upd.fill_synthetic(q, first_patch=5).cpu().numpy()
    want = oracle.fill_synthetic(cfg, 9, dtype=npdt, first_patch=5)
    assert_bitwise(got, want, "synthetic input")


# ----------------------------------------------------------------------------------------- every committed shape
def _committed():
    from exahype_b200 import runtime
    return [(i["model"], i["dim"], i["patch_size"], i["halo_size"], i["n_real"], i["n_aux"], i["dtype"])
            for i in runtime.committed_instantiations()]


@pytest.mark.parametrize("shape", _committed(), ids=lambda s: "-".join(map(str, s)))
@pytest.mark.parametrize("diss", ["var0", "all"])
@pytest.mark.parametrize("output", ["haloed", "unhaloed"])
@pytest.mark.parametrize("kernel", ["auto", "cell"])
def test_every_instantiation_matches_oracle_bitwise(torch, rt, oracle, shape, diss, output, kernel):
    model, dim, P, h, nr, na, dtype = shape
    upd = rt.PatchUpdate(model, dim, P, h, nr, na, dtype=dtype, dissipation=diss, output=output, kernel=kernel)
    if kernel == "cell" and upd.launch_info(10 ** 6) == dataclasses.replace(upd, kernel="auto").launch_info(10 ** 6):
        pytest.skip("this shape has one kernel")
    cfg = oracle_cfg(oracle, upd)
    npdt = np.float64 if dtype == "f64" else np.float32
    info = upd.launch_info(10 ** 6)
    # enough patches that every warp group / CTA streams several patches back to back (ring and staging buffers wrap
    # across patch boundaries), with a ragged tail
    n = max(3 * info["patches_per_tile"] * 5 + 1,
            2 * info["grid"] * info["patches_per_tile"] + 7 if (dim == 3 and info["grid"] < 10 ** 4) else 0)
    q0 = oracle.fill_synthetic(cfg, n, dtype=npdt)
    want = q0.copy()
    lam_o, lmax_o = oracle.step(cfg, want, 0.01, nthreads=4)
    got, lam, lmax = gpu_step(torch, upd, q0, 0.01)
    if output == "haloed":
        assert_bitwise(got, want, "haloed")
    else:
        assert got.shape == upd.out_shape(n)
        assert_bitwise(got, interior(upd, want), "unhaloed")
    assert_bitwise(lam, lam_o, "lambda_patch")
    assert lmax == float(lmax_o)


@pytest.mark.parametrize("n", [1, 7, 8, 9, 147, 148 * 8, 148 * 8 + 1, 3 * 148 * 8 - 5])
@pytest.mark.parametrize("output", ["haloed", "unhaloed"])
def test_warp_per_patch_ragged_batches(torch, rt, oracle, n, output):
    """The 8x8x8 kernel gives every warp its own stream of patches (csrc/fv3d_pair_kernel.cuh): batches that leave
    warps or whole CTAs without a patch, fill the resident warps exactly, or spill one patch into a second round."""
    upd = rt.PatchUpdate("euler", 3, 8, 1, 5, 0, output=output)
    cfg = oracle_cfg(oracle, upd)
    q0 = oracle.fill_synthetic(cfg, n, first_patch=11)
    want = q0.copy()
    lam_o, lmax_o = oracle.step(cfg, want, 0.02, nthreads=4)
    got, lam, lmax = gpu_step(torch, upd, q0, 0.02)
    assert_bitwise(got, want if output == "haloed" else interior(upd, want), output)
    assert_bitwise(lam, lam_o, "lambda_patch")
    assert lmax == float(lmax_o)


def test_fp32_error_bound_against_fp64_oracle(torch, rt, oracle):
    """Stated fp32 bound: max |q32 - q64| <= 2e-6 * max|q64| per variable on the synthetic SWE and Euler states."""
    for model, dim, P, nr, na in (("swe", 2, 32, 3, 1), ("euler", 3, 8, 5, 0), ("euler", 2, 16, 4, 0)):
        upd32 = rt.PatchUpdate(model, dim, P, 1, nr, na, dtype="f32", dissipation="all")
        cfg = oracle_cfg(oracle, upd32)
        q64 = oracle.fill_synthetic(cfg, 64)
        q32 = q64.astype(np.float32)
        oracle.step(cfg, q64, 0.01, nthreads=4)
        got, _, _ = gpu_step(torch, upd32, q32, 0.01)
        scale = np.abs(q64).reshape(-1, nr + na).max(axis=0)
        err = np.abs(got.astype(np.float64) - q64).reshape(-1, nr + na).max(axis=0)
        assert (err <= 2e-6 * scale).all(), (model, err / scale)


@pytest.mark.parametrize("model,P,nr,na,dtype", [("euler", 16, 4, 0, "f64"), ("swe", 32, 3, 1, "f64"), ("swe", 32, 3, 1, "f32")])
def test_row_marching_on_16_byte_aligned_buffers(torch, rt, oracle, model, P, nr, na, dtype):
    """The 2-D row-marching kernel uses 256-bit accesses on 32-byte aligned buffers and falls back to 128-bit accesses
    otherwise (the C ABI only asks for 16-byte alignment): both paths give the same bits."""
    for output in ("haloed", "unhaloed"):
        upd = rt.PatchUpdate(model, 2, P, 1, nr, na, dtype=dtype, output=output, dissipation="all")
        cfg = oracle_cfg(oracle, upd)
        tdt = torch.float64 if dtype == "f64" else torch.float32
        npdt = np.float64 if dtype == "f64" else np.float32
        n = 37
        q0 = oracle.fill_synthetic(cfg, n, dtype=npdt)
        want = q0.copy(); lam_o, lmax_o = oracle.step(cfg, want, 0.01)
        off = 16 // q0.itemsize
        raw_in = torch.zeros(q0.size + off, dtype=tdt, device="cuda")
        q = raw_in[off:].view(q0.shape); q.copy_(torch.from_numpy(q0))
        assert q.data_ptr() % 32 == 16
        n_out = int(np.prod(upd.out_shape(n)))
        raw_out = torch.zeros(n_out + off, dtype=tdt, device="cuda")
        out = raw_out[off:].view(upd.out_shape(n))
        if output == "haloed":
            out.copy_(q)
        lam = torch.zeros(n, dtype=tdt, device="cuda")
        upd.step(q, out, 0.01, lam, None)
        torch.cuda.synchronize()
        assert_bitwise(out.cpu().numpy(), want if output == "haloed" else interior(upd, want), output)
        assert_bitwise(lam.cpu().numpy(), lam_o, "lambda")


def _with_aux():
    return [s for s in _committed() if s[5] > 0 and s[2] in (16, 32)]      # shapes the row-marching kernel serves


@pytest.mark.parametrize("shape", _with_aux(), ids=lambda s: "-".join(map(str, s)))
@pytest.mark.parametrize("diss", ["var0", "all"])
def test_unknowns_only_output_matches_oracle_bitwise(torch, rt, oracle, shape, diss):
    """EXAHYPE_FLAG_OUTPUT_UNKNOWNS_ONLY: q_out[patch][P][P][n_real] holds exactly the unknowns of the un-haloed output
    (the auxiliary variables, which a step never changes, are not repeated), on 32- and 16-byte aligned buffers, for a
    ragged batch; the thread-per-cell kernel and the CellData form have no such output and say so."""
    model, dim, P, h, nr, na, dtype = shape
    upd = rt.PatchUpdate(model, dim, P, h, nr, na, dtype=dtype, dissipation=diss, output="unknowns")
    assert upd.out_shape(5) == (5, P, P, nr)
    cfg = oracle_cfg(oracle, upd)
    npdt = np.float64 if dtype == "f64" else np.float32
    tdt = torch.float64 if dtype == "f64" else torch.float32
    n = 3 * upd.launch_info(10 ** 6)["patches_per_tile"] * 5 + 1
    q0 = oracle.fill_synthetic(cfg, n, dtype=npdt)
    want = q0.copy()
    lam_o, lmax_o = oracle.step(cfg, want, 0.01, nthreads=4)
    got, lam, lmax = gpu_step(torch, upd, q0, 0.01)
    assert got.shape == upd.out_shape(n)
    assert_bitwise(got, interior(upd, want)[..., :nr], "unknowns")
    assert_bitwise(lam, lam_o, "lambda_patch")
    assert lmax == float(lmax_o)
    # 16-byte aligned buffers take the 128-bit loads
    off = 16 // q0.itemsize
    raw_in = torch.zeros(q0.size + off, dtype=tdt, device="cuda")
    q = raw_in[off:].view(q0.shape); q.copy_(torch.from_numpy(q0))
    raw_out = torch.full((int(np.prod(upd.out_shape(n))) + off,), 7.0, dtype=tdt, device="cuda")
    out = raw_out[off:].view(upd.out_shape(n))
    upd.step(q, out, 0.01)
    torch.cuda.synchronize()
    assert_bitwise(out.cpu().numpy(), interior(upd, want)[..., :nr], "unknowns, 16-byte aligned")
    with pytest.raises(rt.ExaHyPECudaError) as e:
        dataclasses.replace(upd, kernel="cell").step(q, out, 0.01)
    assert e.value.code == -2


def test_unknowns_only_output_is_the_unhaloed_one_without_auxiliary_variables(torch, rt, oracle):
    """No auxiliary variables, no second layout: the flag is accepted and changes nothing (3-D Euler, 8^3)."""
    upd = rt.PatchUpdate("euler", 3, 8, 1, 5, 0, output="unknowns")
    cfg = oracle_cfg(oracle, upd)
    q0 = oracle.fill_synthetic(cfg, 41)
    want = q0.copy(); oracle.step(cfg, want, 0.01)
    got, _, _ = gpu_step(torch, upd, q0, 0.01)
    assert_bitwise(got, interior(upd, want), "unknowns == unhaloed")


# densities / water heights far from 1 whose every intermediate stays finite, across 600 orders of magnitude: operands
# for which nvcc's IEEE division / square root leave their fast path (tiny and huge exponents, subnormal reciprocals:
# 1/5e307, a zero radicand) next to ordinary ones
HOSTILE_EULER = [1e-100, -1e-100, 1e-140, 3.0e-151, 3.1e-151, 1e100, 3.2e150, 3.3e150, 1e200, 1e300, 5e307, -1e250]
HOSTILE_SWE = [1e-100, -1e-120, 1e-140, 3.0e-151, 3.1e-151, 1e-152, 1e100, 1e140, 3.2e150, 3.3e150, 1e151]


@pytest.mark.parametrize("model,dim,P,nr,na", [("euler", 3, 8, 5, 0), ("euler", 2, 16, 4, 0), ("swe", 2, 32, 3, 1),
                                               ("euler", 3, 4, 5, 0), ("euler", 2, 3, 4, 0), ("swe_source", 2, 16, 3, 3)])
@pytest.mark.parametrize("diss", ["var0", "all"])
def test_extreme_magnitudes_and_zero_pressure_bit_for_bit(torch, rt, oracle, model, dim, P, nr, na, diss):
    """Bitwise parity away from the benchmark's O(1) states: densities of either sign across 600 orders of magnitude
    (including reciprocals that are subnormal) and exactly zero pressures (sqrt(0)), one of each per patch, anywhere in
    the haloed patch.  These are the operands for which the IEEE division / root take their slow paths; any arithmetic
    policy that is to replace the per-operation IEEE code (the branch-free ArithFast sequences are bit-identical only
    inside the fast paths' range) has to pass this.  (Inputs that drive the REFERENCE to infinities or
    NaNs are outside what this checks: the cheaper physics forms of csrc/physics.cuh are identities for finite values.)"""
    upd = rt.PatchUpdate(model, dim, P, 1, nr, na, dissipation=diss)
    cfg = oracle_cfg(oracle, upd)
    n = 3 * upd.launch_info(10 ** 6)["patches_per_tile"] * 3 + 5
    q0 = oracle.fill_synthetic(cfg, n)
    rng = np.random.default_rng(3)
    hostile = HOSTILE_EULER if model == "euler" else HOSTILE_SWE
    side = P + 2
    for p_ in range(n):        # ONE hostile density per patch (two of them side by side would overflow the reference
        a = rng.integers(0, side, dim)                   # itself) and one zero-pressure cell
        q0[(p_,) + tuple(a) + (0,)] = hostile[p_ % len(hostile)]
        if model == "euler":
            b = (a + 1 + rng.integers(0, side - 1, dim)) % side        # differs from a in every index
            q0[(p_,) + tuple(b) + (slice(1, dim + 2),)] = 0.0          # no momentum, no energy -> p = 0 -> sqrt(0)
    want = q0.copy()
    lam_o, lmax_o = oracle.step(cfg, want, 0.01, nthreads=4)
    assert np.isfinite(want).all() and np.isfinite(lam_o).all(), "the test's states must keep the reference finite"
    got, lam, lmax = gpu_step(torch, upd, q0, 0.01)
    assert_bitwise(got, want, "state")
    assert_bitwise(lam, lam_o, "lambda_patch")
    assert lmax == float(lmax_o)


# ----------------------------------------------------------------------------------------- semantics of the boundary
def test_out_of_place_haloed_writes_interior_only(torch, rt, oracle):
    upd = rt.PatchUpdate("euler", 3, 8, 1, 5, 0)
    cfg = oracle_cfg(oracle, upd)
    q0 = oracle.fill_synthetic(cfg, 37)
    q = torch.from_numpy(q0).cuda()
    out = torch.full_like(q, -123.0)
    upd.step(q, out, 0.01)
    torch.cuda.synchronize()
    out = out.cpu().numpy()
    assert_bitwise(q.cpu().numpy(), q0, "input untouched")
    want = q0.copy(); oracle.step(cfg, want, 0.01)
    assert_bitwise(interior(upd, out), interior(upd, want), "interior")
    mask = np.ones(out.shape, bool); mask[:, 1:-1, 1:-1, 1:-1] = False
    assert (out[mask] == -123.0).all()


def test_lambda_accumulates_across_launches_and_empty_batch(torch, rt, oracle):
    upd = rt.PatchUpdate("euler", 2, 16, 1, 4, 0)
    cfg = oracle_cfg(oracle, upd)
    q0 = oracle.fill_synthetic(cfg, 64)
    lam_o, lmax_o = oracle.step(cfg, q0.copy(), 0.01)
    q = torch.from_numpy(q0).cuda()
    lmax = torch.zeros(1, dtype=torch.float64, device="cuda")
    half = q0.shape[0] // 2
    upd.step(q[:half].clone(), None, 0.01, None, lmax)
    upd.step(q[half:].clone(), None, 0.01, None, lmax, accumulate_lambda=True)
    assert float(lmax.item()) == lmax_o
    upd.step(q[:0], None, 0.01, None, lmax)          # zero patches: resets lambda_max, launches nothing
    assert float(lmax.item()) == 0.0


def test_host_time_step_equals_device_step(torch, rt, oracle):
    """The reference's call shape time_step(Q, dt) on host memory, chunked so several pipeline slots are used."""
    upd = rt.PatchUpdate("euler", 3, 8, 1, 5, 0)
    cfg = oracle_cfg(oracle, upd)
    q0 = oracle.fill_synthetic(cfg, 301)
    want = q0.copy(); lam_o, lmax_o = oracle.step(cfg, want, 0.01, nthreads=4)
    rt.load().exahype_cuda_host_pipeline_configure(64, 3)
    try:
        Q = q0.copy()
        lam = np.zeros(301)
        lmax = upd.time_step(Q, 0.01, lambda_patch=lam)
        assert_bitwise(Q, want, "host in place")
        assert_bitwise(lam, lam_o, "host lambda")
        assert lmax == lmax_o
        upd_u = dataclasses.replace(upd, output="unhaloed")
        out = np.zeros(upd_u.out_shape(301))
        upd_u.time_step(q0.copy(), 0.01, Q_out=out)
        assert_bitwise(out, interior(upd, want), "host unhaloed")
        pinned = torch.from_numpy(q0.copy()).pin_memory()
        upd.time_step(pinned.numpy(), 0.01)
        assert_bitwise(pinned.numpy(), want, "pinned")
    finally:
        rt.load().exahype_cuda_host_pipeline_configure(0, 3)
        rt.load().exahype_cuda_host_pipeline_release()


def test_errors_on_device(torch, rt):
    upd = rt.PatchUpdate("euler", 3, 8, 1, 5, 0)
    q = torch.zeros(upd.in_shape(2), dtype=torch.float64, device="cuda")
    with pytest.raises(ValueError):
        upd.step(q.float(), None, 0.1)
    with pytest.raises(ValueError):
        upd.step(torch.zeros(17, dtype=torch.float64, device="cuda"), None, 0.1)
    with pytest.raises(rt.ExaHyPECudaError) as e:
        rt.PatchUpdate("euler", 3, 7, 1, 5, 0).step(torch.zeros((1, 9, 9, 9, 5), dtype=torch.float64, device="cuda"), None, 0.1)
    assert e.value.code == -2
    before = rt.launch_count()
    upd.step(q + 1.0, None, 0.1)
    torch.cuda.synchronize()
    assert rt.launch_count() == before + 1


# ----------------------------------------------------------------------------------------- CellData form
@pytest.mark.parametrize("model,dim,P,nr,na,dtype", [
    ("euler", 3, 8, 5, 0, "f64"), ("euler", 3, 4, 5, 0, "f32"), ("euler", 2, 16, 4, 0, "f64"), ("euler", 2, 3, 4, 0, "f64"),
    ("euler", 2, 4, 5, 5, "f64"), ("swe", 2, 32, 3, 1, "f32"), ("swe", 2, 32, 3, 1, "f64")])
@pytest.mark.parametrize("output", ["haloed", "unhaloed"])
def test_cell_data_form_gathered_patches_and_per_patch_dt(torch, rt, oracle, model, dim, P, nr, na, dtype, output):
    """exahype_cuda_fv_step_cell_data: patches scattered through a pool in permuted order (QIn[p] / QOut[p] pointers),
    each with its own dt (CellData::dt), per-patch eigenvalue into CellData::maxEigenvalue.  Every patch must equal the
    oracle run on that patch alone with that dt, bit for bit; the pool outside the addressed patches stays untouched."""
    upd = rt.PatchUpdate(model, dim, P, 1, nr, na, dtype=dtype, output=output, dissipation="all")
    cfg = oracle_cfg(oracle, upd)
    tdt = torch.float64 if dtype == "f64" else torch.float32
    npdt = np.float64 if dtype == "f64" else np.float32
    n, pool_n = 45, 64
    rng = np.random.default_rng(7)
    slots = rng.permutation(pool_n)[:n]                     # patch p lives in pool slot slots[p]
    dts = rng.uniform(0.002, 0.02, n).astype(npdt)
    q0 = oracle.fill_synthetic(cfg, n, dtype=npdt)
    pool = torch.full(upd.in_shape(pool_n), -5.0, dtype=tdt, device="cuda")
    pool[torch.from_numpy(slots).cuda()] = torch.from_numpy(q0).cuda()
    per_in = int(np.prod(upd.in_shape(1))) * pool.element_size()
    in_ptrs = torch.tensor([pool.data_ptr() + int(s) * per_in for s in slots], dtype=torch.int64, device="cuda")
    if output == "haloed":
        out_pool, out_ptrs = pool, in_ptrs                  # in place, the reference's semantics
    else:
        out_pool = torch.full(upd.out_shape(pool_n), -7.0, dtype=tdt, device="cuda")
        per_out = int(np.prod(upd.out_shape(1))) * out_pool.element_size()
        out_slots = rng.permutation(pool_n)[:n]
        out_ptrs = torch.tensor([out_pool.data_ptr() + int(s) * per_out for s in out_slots], dtype=torch.int64, device="cuda")
    lam = torch.zeros(n, dtype=tdt, device="cuda")
    lmax = torch.zeros(1, dtype=tdt, device="cuda")
    upd.step_cell_data(in_ptrs, out_ptrs, dt_patch=torch.from_numpy(dts).cuda(), max_eigenvalue=lam, lambda_max=lmax)
    torch.cuda.synchronize()
    want = q0.copy()
    lam_o = np.zeros(n, dtype=npdt)
    for p in range(n):
        one = want[p:p + 1]
        l, _ = oracle.step(cfg, one, float(dts[p]))
        lam_o[p] = l[0]
    assert_bitwise(lam.cpu().numpy(), lam_o, "maxEigenvalue")
    assert float(lmax.item()) == float(lam_o.max())
    out_np = out_pool.cpu().numpy()
    if output == "haloed":
        assert_bitwise(out_np[slots], want, "in-place patches")
        untouched = np.setdiff1d(np.arange(pool_n), slots)
        assert (out_np[untouched] == -5.0).all()
    else:
        assert_bitwise(out_np[out_slots], interior(upd, want), "QOut patches")
        untouched = np.setdiff1d(np.arange(pool_n), out_slots)
        assert (out_np[untouched] == -7.0).all()
        assert_bitwise(pool.cpu().numpy()[slots], q0, "QIn untouched")


@pytest.mark.parametrize("output,dissipation", [("unhaloed", "var0"), ("haloed", "var0"), ("unhaloed", "all")])
def test_cell_data_form_several_patches_per_warp(torch, rt, oracle, output, dissipation):
    """The warp-per-patch kernel in its CellData form with more patches than resident warps (148 SMs x 8): every warp
    walks through three or four patches and reads the table entries (QIn, QOut, dt) of its next patch one patch ahead.
    Permuted pointers, per-patch dt, a ragged count; every patch bit for bit the oracle's result for its own dt."""
    upd = rt.PatchUpdate("euler", 3, 8, 1, 5, 0, output=output, dissipation=dissipation)
    cfg = oracle_cfg(oracle, upd)
    n, pool_n = 4096 + 37, 4096 + 37 + 11
    rng = np.random.default_rng(11)
    slots = rng.permutation(pool_n)[:n]
    dt_values = np.array([0.004, 0.01, 0.017])
    which = rng.integers(0, 3, n)
    dts = dt_values[which]
    q0 = oracle.fill_synthetic(cfg, n)
    pool = torch.full(upd.in_shape(pool_n), -5.0, dtype=torch.float64, device="cuda")
    pool[torch.from_numpy(slots).cuda()] = torch.from_numpy(q0).cuda()
    per_in = int(np.prod(upd.in_shape(1))) * 8
    in_ptrs = torch.from_numpy(pool.data_ptr() + slots.astype(np.int64) * per_in).cuda()
    if output == "haloed":
        out_pool, out_ptrs, out_slots = pool, in_ptrs, slots
    else:
        out_pool = torch.full(upd.out_shape(pool_n), -7.0, dtype=torch.float64, device="cuda")
        per_out = int(np.prod(upd.out_shape(1))) * 8
        out_slots = rng.permutation(pool_n)[:n]
        out_ptrs = torch.from_numpy(out_pool.data_ptr() + out_slots.astype(np.int64) * per_out).cuda()
    lam = torch.zeros(n, dtype=torch.float64, device="cuda")
    lmax = torch.zeros(1, dtype=torch.float64, device="cuda")
    upd.step_cell_data(in_ptrs, out_ptrs, dt_patch=torch.from_numpy(dts).cuda(), max_eigenvalue=lam, lambda_max=lmax)
    torch.cuda.synchronize()
    want = q0.copy()
    lam_o = np.zeros(n)
    for g, dt in enumerate(dt_values):                     # the oracle once per distinct dt
        idx = np.nonzero(which == g)[0]
        part = np.ascontiguousarray(want[idx])
        l, _ = oracle.step(cfg, part, float(dt), nthreads=4)
        want[idx] = part
        lam_o[idx] = l
    assert_bitwise(lam.cpu().numpy(), lam_o, "maxEigenvalue")
    assert float(lmax.item()) == float(lam_o.max())
    out_np = out_pool.cpu().numpy()
    if output == "haloed":
        assert_bitwise(out_np[slots], want, "in-place patches")
    else:
        assert_bitwise(out_np[out_slots], interior(upd, want), "QOut patches")
        assert_bitwise(pool.cpu().numpy()[slots], q0, "QIn untouched")
    untouched = np.setdiff1d(np.arange(pool_n), out_slots)
    assert (out_np[untouched] == (-5.0 if output == "haloed" else -7.0)).all()


@pytest.mark.parametrize("model,P,nr,na", [("euler", 16, 4, 0), ("swe", 32, 3, 1), ("euler", 8, 4, 0)])
@pytest.mark.parametrize("output", ["haloed", "unhaloed"])
def test_cell_data_form_mixed_patch_alignment(torch, rt, oracle, model, P, nr, na, output):
    """The gathered row-marching kernel picks 256-bit or 128-bit accesses per lane from the alignment of that lane's
    patch: a pool whose slots alternate between 32-byte and 16-byte alignment, permuted, so that the patches of one warp
    differ -- every patch still equals the oracle bit for bit and nothing else in the pool changes."""
    upd = rt.PatchUpdate(model, 2, P, 1, nr, na, output=output)
    cfg = oracle_cfg(oracle, upd)
    n, pool_n = 75, 96
    rng = np.random.default_rng(11)
    slots = rng.permutation(pool_n)[:n]
    q0 = oracle.fill_synthetic(cfg, n)
    per_in = int(np.prod(upd.in_shape(1)))
    stride_in = per_in + 2                                   # 16 bytes of padding: odd slots are 16-byte aligned only
    pool = torch.full((pool_n * stride_in,), -5.0, dtype=torch.float64, device="cuda")
    assert pool.data_ptr() % 32 == 0 and (per_in * 8) % 32 == 0
    for p, s_ in enumerate(slots):
        pool[s_ * stride_in:s_ * stride_in + per_in] = torch.from_numpy(q0[p].ravel()).cuda()
    in_ptrs = torch.tensor([pool.data_ptr() + int(s_) * stride_in * 8 for s_ in slots], dtype=torch.int64, device="cuda")
    assert len({int(x) % 32 for x in in_ptrs.tolist()}) == 2
    if output == "haloed":
        out_pool, out_ptrs, per_out, stride_out, out_slots = pool, in_ptrs, per_in, stride_in, slots
    else:
        per_out = int(np.prod(upd.out_shape(1)))
        stride_out = per_out + 2
        out_slots = rng.permutation(pool_n)[:n]
        out_pool = torch.full((pool_n * stride_out,), -7.0, dtype=torch.float64, device="cuda")
        out_ptrs = torch.tensor([out_pool.data_ptr() + int(s_) * stride_out * 8 for s_ in out_slots], dtype=torch.int64, device="cuda")
    lam = torch.zeros(n, dtype=torch.float64, device="cuda")
    upd.step_cell_data(in_ptrs, out_ptrs, dt=0.01, max_eigenvalue=lam)
    torch.cuda.synchronize()
    want = q0.copy()
    lam_o, _ = oracle.step(cfg, want, 0.01, nthreads=4)
    assert_bitwise(lam.cpu().numpy(), lam_o, "maxEigenvalue")
    out_np = out_pool.cpu().numpy().reshape(pool_n, stride_out)
    fill = -5.0 if output == "haloed" else -7.0
    assert (out_np[:, per_out:] == fill).all(), "padding between the slots"
    expect = want if output == "haloed" else interior(upd, want)
    assert_bitwise(out_np[out_slots, :per_out].reshape(expect.shape), expect, "gathered patches")
    untouched = np.setdiff1d(np.arange(pool_n), out_slots)
    assert (out_np[untouched] == fill).all()


# ----------------------------------------------------------------------------------------- BASELINE sizes
@pytest.mark.parametrize("model,dim,P,nr,na,n,dtype", [
    ("euler", 3, 8, 5, 0, 32768, "f64"), ("euler", 2, 16, 4, 0, 65536, "f64"),
    ("swe", 2, 32, 3, 1, 65536, "f64"), ("swe", 2, 32, 3, 1, 65536, "f32")])      # C3, C2, C4, C4 fp32 as benched
def test_full_size_properties(torch, rt, oracle, model, dim, P, nr, na, n, dtype):
    """At the full BASELINE batch sizes: determinism, shard-independence (what multi-GPU relies on), untouched
    halos, lambda_max == max(lambda_patch), and bitwise parity on the first / last / middle patches (fp32 against
    the fp32 oracle)."""
    upd = rt.PatchUpdate(model, dim, P, 1, nr, na, dtype=dtype)
    cfg = oracle_cfg(oracle, upd)
    tdt = torch.float64 if dtype == "f64" else torch.float32
    q0 = oracle.fill_synthetic(cfg, n, dtype=np.float64 if dtype == "f64" else np.float32)
    q = torch.from_numpy(q0).cuda()
    lam = torch.zeros(n, dtype=tdt, device="cuda")
    lmax = torch.zeros(1, dtype=tdt, device="cuda")
    a = q.clone(); upd.step(a, None, 0.01, lam, lmax)
    b = q.clone(); upd.step(b, None, 0.01)
    assert torch.equal(a, b)
    # two shards processed separately == the whole batch
    c = q.clone(); cut = n // 2 + 3
    upd.step(c[:cut], None, 0.01); upd.step(c[cut:], None, 0.01)
    assert torch.equal(a, c)
    assert float(lmax.item()) == float(lam.max().item())
    a_np = a.cpu().numpy()
    mask = np.ones(a_np.shape[1:], bool); mask[(slice(1, -1),) * dim] = False
    assert_bitwise(a_np[:, mask], q0[:, mask], "halos")
    for lo in (0, n // 2 - 32, n - 64):
        want = q0[lo:lo + 64].copy()
        lam_o, _ = oracle.step(cfg, want, 0.01, nthreads=4)
        assert_bitwise(a_np[lo:lo + 64], want, f"patches {lo}..")
        assert_bitwise(lam[lo:lo + 64].cpu().numpy(), lam_o, "lambda")
