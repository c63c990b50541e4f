"""The device functors evaluate two quantities of the reference's physics by cheaper, algebraically different
expressions (exahype_b200/csrc/physics.cuh) and claim the SAME BITS:

    max(|u - c|, |u + c|),  u = q / |rho|     (Unit test/Functions.cpp:56-58)   ==   |w + copysign(c, w)|,  w = (1/rho) * q
    0.5 * |1/rho| * ke                        (Functions.cpp:50-51)             ==   |0.5 * (1/rho) * ke|

The GPU parity tests check this end to end against the oracle; this test pins the identities themselves on the CPU in
IEEE double and single precision (numpy: round-to-nearest-even, no contraction), on random and on hostile inputs."""
import numpy as np
import pytest


def _samples(dtype, n=400_000, seed=7):
    rng = np.random.default_rng(seed)
    fi = np.finfo(dtype)
    mag = np.exp(rng.uniform(np.log(fi.tiny * 4), np.log(fi.max ** 0.25), n)).astype(dtype)
    sign = rng.choice(np.array([-1, 1], dtype=dtype), n)
    x = mag * sign
    special = np.array([0.0, -0.0, fi.tiny, -fi.tiny, fi.smallest_subnormal, -fi.smallest_subnormal, 1.0, -1.0,
                        1.0 + fi.eps, -(1.0 + fi.eps), 3.0, 1e-3, fi.max ** 0.25], dtype=dtype)
    return np.concatenate([x, special, rng.permutation(np.resize(special, 997))])


def _bits(a):
    return a.view(np.uint64 if a.dtype == np.float64 else np.uint32)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_wave_speed_identity_is_bitwise(dtype):
    rho = _samples(dtype, seed=1)
    q = _samples(dtype, seed=2)[::-1].copy()
    c = np.abs(_samples(dtype, seed=3))                       # a sound speed is >= +0
    with np.errstate(divide="ignore", over="ignore", invalid="ignore", under="ignore"):
        irho_abs = dtype(1.0) / np.abs(rho)                    # Functions.cpp:50
        u = q * irho_abs                                       # Functions.cpp:56
        want = np.maximum(np.abs(u - c), np.abs(u + c))        # std::max on non-NaN values
        irho = dtype(1.0) / rho                                # Functions.cpp:20 (the flux's)
        w = irho * q
        got = np.abs(w + np.copysign(c, w))
    ok = np.isfinite(want) & np.isfinite(got)
    assert ok.sum() > 0.9 * ok.size
    assert np.array_equal(_bits(got[ok]), _bits(want[ok]))
    assert np.array_equal(np.isfinite(want), np.isfinite(got))


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_pressure_from_absolute_density_identity_is_bitwise(dtype):
    rho = _samples(dtype, seed=4)
    ke = np.abs(_samples(dtype, seed=5))                       # a sum of squares
    with np.errstate(divide="ignore", over="ignore", invalid="ignore", under="ignore"):
        want = dtype(0.5) * (dtype(1.0) / np.abs(rho)) * ke
        got = np.abs(dtype(0.5) * (dtype(1.0) / rho) * ke)
    ok = np.isfinite(want)
    assert np.array_equal(_bits(got[ok]), _bits(want[ok]))
