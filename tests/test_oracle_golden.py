"""Pins the CPU oracle (test infrastructure) before anything is compared against it.

Sources of truth, all produced by the reference itself or recorded in SURVEY.md section 8c:
* the reference's committed generated kernel compiled unmodified (oracle/_ref, dev container) and its
  committed output fixture tests/golden/g0_reference_kernel.json (travels to the GPU box);
* goldens G0h / G1 / G3s / G3r / G2r (hashes and spot values at rtol 1e-12).
"""
import dataclasses
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _bits(hexes):
    return np.array([int(h, 16) for h in hexes], dtype=np.uint64).view(np.float64)


@pytest.fixture(scope="module")
def g0():
    with open(os.path.join(HERE, "golden", "g0_reference_kernel.json")) as f:
        return json.load(f)


def test_g0_fixture_is_the_survey_golden(g0):
    assert g0["fnv1a64"] == "c192d87bfc9efcb9"
    out = _bits(g0["output"]).reshape(6, 6, 10)
    np.testing.assert_allclose(out[1, 1, :5], [-0.70557128425959426, -0.20750500759389573, -0.20557763105551741,
                                               -0.20361748135538105, 0.60171772632281284], rtol=1e-15)
    np.testing.assert_allclose(out[2, 2, :5], [0.53869414698947504, 0.75203086118104856, 0.75980698279961478,
                                               0.76748154968333859, 0.95098323355963421], rtol=1e-15)


def test_oracle_committed_ranges_reproduce_reference_kernel_bitwise(oracle, g0):
    cfg = oracle.REFERENCE_CONFIG
    q = oracle.fill_sin(cfg, 1)
    assert np.array_equal(q.reshape(-1), _bits(g0["input"]))
    oracle.step(cfg, q, g0["dt"])
    assert np.array_equal(q.reshape(-1).view(np.uint64), _bits(g0["output"]).view(np.uint64))
    assert oracle.fnv1a64(q) == g0["fnv1a64"]


def test_compiled_reference_matches_fixture(oracle, g0):
    if not oracle.reference_available():
        pytest.skip("oracle/_ref not built (needs /root/reference; dev container only)")
    q = _bits(g0["input"]).copy()
    oracle.reference_time_step(q, g0["dt"])
    assert np.array_equal(q.view(np.uint64), _bits(g0["output"]).view(np.uint64))


def test_oracle_physics_matches_reference_functions(oracle, g0):
    """Flux / maxEigenvalue of Functions.cpp on random admissible states, through a 1-patch step is indirect;
    here the samples recorded from the reference's own functions are replayed against a tiny oracle run."""
    cfg = oracle.OracleConfig(dim=2, patch_size=1, halo=1, n_real=4, n_aux=0)
    for s in g0["physics_samples"]:
        q = _bits(s["q"])[:4]
        n = s["normal"]
        # centre cell holds q, all neighbours hold q too -> lambda_patch == maxEigenvalue(q, n) maximised over n
        Q = np.tile(q, (1, 3, 3, 1)).copy()
        lam, _ = oracle.step(cfg, Q, 0.0)
        lam_ref = [float(_bits([t["lambda"]])[0]) for t in g0["physics_samples"] if t["q"] == s["q"]]
        assert lam[0] == max(lam_ref)
        # flux: with q on the -n side replaced by zero momentum/energy scaling we would need F itself; check F via
        # the update formula Qc = Q - 0.5 F(right) + 0.5 F(left) on a patch whose right neighbour along n is q
        # and whose left neighbour has F == 0 is not constructible for Euler, so compare F directly instead:
        F_ref = _bits(s["F"])[:4]
        Q = np.tile(q, (1, 3, 3, 1)).copy()
        q_left = q.copy(); q_left[1:3] = 0.0          # zero momentum: F = p * e_n only
        idx = [0, 1, 1]; idx[1 + n] = 0
        Q[tuple(idx)] = q_left
        centre = Q[0, 1, 1].copy()
        oracle.step(cfg, Q, 0.0)
        irho = 1.0 / q_left[0]
        p_left = (1.4 - 1) * (q_left[3] - 0.5 * irho * 0.0)
        F_left = np.zeros(4); F_left[n + 1] += p_left
        # other axis: both neighbours equal -> (c - 0.5 F) + 0.5 F, reproduced with the reference's F values
        expect = centre.copy()
        for axis in (0, 1):
            Fp = _bits([t["F"] for t in g0["physics_samples"] if t["q"] == s["q"] and t["normal"] == axis][0])[:4]
            Fm = F_left if axis == n else Fp
            expect = expect - 0.5 * Fp + 0.5 * Fm
        assert np.array_equal(Q[0, 1, 1], expect), (n, F_ref)


G1 = dict(dim=2, patch_size=3, halo=1, n_real=4, n_aux=0)
G3 = dict(dim=3, patch_size=8, halo=1, n_real=5, n_aux=0)
G2 = dict(dim=2, patch_size=16, halo=1, n_real=4, n_aux=0)


@pytest.mark.parametrize("diss,dt,expected", [(0, 1.0, "1334ede917c43f96"), (1, 1.0, "aa5c6cb6b3c1c18e"),
                                              (0, 0.01, "fbacf0e93e60875b"), (1, 0.01, "369e22f97aff515c")])
def test_g1_sin_input_config_c1(oracle, diss, dt, expected):
    cfg = oracle.OracleConfig(diss=diss, **G1)
    q = oracle.fill_sin(cfg, 1000)
    halo_before = q.copy()
    lam, lmax = oracle.step(cfg, q, dt)
    assert oracle.fnv1a64(q) == expected
    if diss == 0 and dt == 1.0:
        np.testing.assert_allclose(q[0, 1, 1], [-0.00097217991318677955, 3.1858929737296868e-05,
                                                7.0631932862739674e-05, 0.00010147874503622743], rtol=1e-12)
        np.testing.assert_allclose(q[499, 2, 2], [0.99999525652023169, 0.9999967950076063, 0.99999687781597046,
                                                  0.99999695963774726], rtol=1e-12)
        np.testing.assert_allclose(q[999, 3, 3], [0.0029810776535199725, 0.0021944470744585756,
                                                  0.0021622802172080978, 0.0021307318517740149], rtol=1e-12)
    np.testing.assert_allclose([lam[0], lam[499], lam[999], lmax],
                               [1.1326339818690889, 1.0000334079718405, 1.0039091839698087, 1.1326339818690889],
                               rtol=1e-14)
    # halos are never written
    mask = np.ones(q.shape, bool); mask[:, 1:-1, 1:-1] = False
    assert np.array_equal(q[mask], halo_before[mask])


def test_g0h_head_ranges(oracle):
    cfg = dataclasses.replace(oracle.REFERENCE_CONFIG, ranges=oracle.RANGES_HEAD)
    q = oracle.fill_sin(cfg, 1)
    aux_before = q[..., 5:].copy()
    oracle.step(cfg, q, 1.0)
    assert oracle.fnv1a64(q) == "5c1d83d32c1d0f28"
    assert np.array_equal(q[..., 5:], aux_before)          # aux pass through
    committed = oracle.fill_sin(oracle.REFERENCE_CONFIG, 1)
    oracle.step(oracle.REFERENCE_CONFIG, committed, 1.0)
    assert np.array_equal(q[0, 2:4, 2:4], committed[0, 2:4, 2:4])   # unaffected by the uninitialised rows


@pytest.mark.parametrize("diss,expected,cell", [
    (0, "53201509629d2991", [-0.001041178244064389, 0.016086051093805571, 0.0021848184460915993,
                             0.0009359514350143136, 0.018282559964790283]),
    (1, "78219e9b15eba6f7", [-0.001041178244064389, 0.014742553746402143, 0.00084133950242203386,
                             -0.00040750907178488177, 0.016939117927997512])])
def test_g3s_euler3d_sin(oracle, diss, expected, cell):
    cfg = oracle.OracleConfig(diss=diss, **G3)
    q = oracle.fill_sin(cfg, 4)
    lam, lmax = oracle.step(cfg, q, 0.01)
    assert oracle.fnv1a64(q) == expected
    np.testing.assert_allclose(q[0, 1, 1, 1], cell, rtol=1e-12)
    np.testing.assert_allclose(lam, [1.5364524048789425, 1.5296852891027781, 1.5291221124861072,
                                     1.5287655424826991], rtol=1e-14)
    assert lmax == lam[0]


@pytest.mark.parametrize("diss,expected", [(0, "f8486ef10dfe78d6"), (1, "c72ba3f4ebf895ec")])
def test_g3r_euler3d_synthetic(oracle, diss, expected):
    cfg = oracle.OracleConfig(diss=diss, **G3)
    q = oracle.fill_synthetic(cfg, 4)
    assert oracle.fnv1a64(q) == "f52d7d23f347b4b8"
    np.testing.assert_allclose(q[0, 0, 0, 0], [1.7131115268902648, 0.38152489638473408, 0.22564911683226299,
                                               0.77545445577354355, 4.7589498765395293], rtol=1e-15)
    lam, lmax = oracle.step(cfg, q, 0.01)
    assert oracle.fnv1a64(q) == expected
    np.testing.assert_allclose(lam, [2.119555658951096, 2.1234716419928956, 2.0686209688018007,
                                     2.0659179394511646], rtol=1e-14)
    assert lmax == lam[1]


@pytest.mark.parametrize("diss,expected", [(0, "1856adbc23fec8cd"), (1, "8e53779668580e14")])
def test_g2r_euler2d_synthetic(oracle, diss, expected):
    cfg = oracle.OracleConfig(diss=diss, **G2)
    q = oracle.fill_synthetic(cfg, 8)
    assert oracle.fnv1a64(q) == "c911a40079f5ebc8"
    lam, lmax = oracle.step(cfg, q, 0.01)
    assert oracle.fnv1a64(q) == expected
    np.testing.assert_allclose(lam[:4], [2.1042037334449715, 2.1332213875447144, 2.0778633934426889,
                                         2.0984832356156726], rtol=1e-14)
    assert lmax == 2.1332213875447144


def test_threads_and_shards_are_bitwise_identical(oracle):
    cfg = oracle.OracleConfig(**G3)
    q1 = oracle.fill_synthetic(cfg, 16)
    q2 = q1.copy()
    lam1, m1 = oracle.step(cfg, q1, 0.01, nthreads=1)
    lam2, m2 = oracle.step(cfg, q2, 0.01, nthreads=4)
    assert np.array_equal(q1, q2) and np.array_equal(lam1, lam2) and m1 == m2
    # a shard generated on its own equals the slice of the whole batch (counter-based input)
    shard = oracle.fill_synthetic(cfg, 4, first_patch=8)
    whole = oracle.fill_synthetic(cfg, 16)
    assert np.array_equal(shard, whole[8:12])


def test_fp32_oracle_tracks_fp64(oracle):
    cfg = oracle.OracleConfig(dim=2, patch_size=32, halo=1, n_real=3, n_aux=1, model=oracle.MODEL_SWE)
    q64 = oracle.fill_synthetic(cfg, 4)
    q32 = oracle.fill_synthetic(cfg, 4, dtype=np.float32)
    assert np.array_equal(q32, q64.astype(np.float32))
    bathy = q64[..., 3].copy()
    oracle.step(cfg, q64, 0.01)
    oracle.step(cfg, q32, 0.01)
    assert np.array_equal(q64[..., 3], bathy)
    # the ONE stated fp32 bound (tests/test_gpu_parity.py): max-norm, per variable,
    # max |q32 - q64| <= 2e-6 * max |q64|
    scale = np.abs(q64).reshape(-1, 4).max(axis=0)
    err = np.abs(q32.astype(np.float64) - q64).reshape(-1, 4).max(axis=0)
    assert (err <= 2e-6 * scale).all(), err / scale


def test_rejects_invalid_configurations(oracle):
    q = np.zeros(10)
    for bad in (dict(dim=4, patch_size=3), dict(dim=2, patch_size=0), dict(dim=2, patch_size=3, halo=0),
                dict(dim=3, patch_size=3, n_real=4)):
        with pytest.raises(ValueError):
            oracle.step(oracle.OracleConfig(**bad), np.zeros(oracle.OracleConfig(**bad).values_per_patch or 1), 1.0)
