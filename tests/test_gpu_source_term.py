"""Source term on the GPU: the committed shallow-water-with-bathymetry-source instantiations
(EXAHYPE_MODEL_SWE_SOURCE: row marching for 32x32 / 16x16 patches, thread per cell for 4x4) and a kernel generated from
the declaration with the user's own device source, all bit for bit against the oracle (fp64 and fp32), both output
forms, both dissipation variants, with guard bands around every buffer."""
import numpy as np
import pytest

import source_term_common as S

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch():
    import torch
    assert torch.cuda.is_available()
    return torch


@pytest.fixture(scope="module")
def rt():
    from exahype_b200 import runtime
    return runtime


def ocfg(oracle, P, diss="var0"):
    return oracle.OracleConfig(dim=2, patch_size=P, halo=1, n_real=3, n_aux=3, model=oracle.MODEL_SWE_SOURCE,
                               diss=oracle.DISS_ALL if diss == "all" else oracle.DISS_VAR0)


@pytest.mark.parametrize("P,dtype,B", [(32, "f64", 150), (32, "f32", 150), (16, "f64", 333), (4, "f64", 1000)])
@pytest.mark.parametrize("diss", ["var0", "all"])
@pytest.mark.parametrize("output", ["haloed", "unhaloed"])
def test_committed_source_instantiations_equal_oracle_bitwise(torch, rt, oracle, P, dtype, B, diss, output):
    upd = rt.PatchUpdate("swe_source", 2, P, 1, 3, 3, dtype=dtype, dissipation=diss, output=output)
    assert upd.supported()
    npdt = np.float64 if dtype == "f64" else np.float32
    q0 = oracle.fill_synthetic(ocfg(oracle, P, diss), B, dtype=npdt)
    want = q0.copy()
    lam_o, lmax_o = oracle.step(ocfg(oracle, P, diss), want, 0.01, nthreads=4)
    guard = 256
    raw = torch.full((q0.size + 2 * guard,), -5.0, dtype=torch.from_numpy(q0).dtype, device="cuda")
    q = raw[guard:guard + q0.size].view(q0.shape)
    q.copy_(torch.from_numpy(q0))
    out = q if output == "haloed" else torch.full(upd.out_shape(B), 9.0, dtype=q.dtype, device="cuda")
    lam = torch.zeros(B, dtype=q.dtype, device="cuda")
    lmax = torch.zeros(1, dtype=q.dtype, device="cuda")
    upd.step(q, out, 0.01, lam, lmax)
    torch.cuda.synchronize()
    assert bool((raw[:guard] == -5.0).all()) and bool((raw[-guard:] == -5.0).all())
    got = out.cpu().numpy()
    if output == "unhaloed":
        want = want[:, 1:-1, 1:-1, :]
    assert np.array_equal(got, want)
    assert np.array_equal(lam.cpu().numpy(), lam_o) and float(lmax.item()) == float(lmax_o)
    # the source changed the momenta: the same batch without it differs
    plain = oracle.OracleConfig(dim=2, patch_size=P, halo=1, n_real=3, n_aux=3, model=oracle.MODEL_SWE,
                                diss=oracle.DISS_ALL if diss == "all" else oracle.DISS_VAR0)
    ref = q0.copy()
    oracle.step(plain, ref, 0.01, nthreads=4)
    assert not np.array_equal(ref if output == "haloed" else ref[:, 1:-1, 1:-1, :], got)


@pytest.mark.parametrize("P", [8, 16])
def test_generated_kernel_with_user_source_term_equals_oracle(torch, oracle, tmp_path, P):
    from exahype.printers import CUDAPrinter
    gk = CUDAPrinter(S.declare(patch_size=P, device_source=True), function_name=f"swe_source_{P}").build(directory=str(tmp_path))
    B = 200
    q0 = oracle.fill_synthetic(ocfg(oracle, P), B)
    want = q0.copy()
    lam_o, lmax_o = oracle.step(ocfg(oracle, P), want, 0.01, nthreads=4)
    q = torch.from_numpy(q0).cuda()
    lam = torch.zeros(B, dtype=torch.float64, device="cuda")
    gk.step(q, None, 0.01, lam)
    torch.cuda.synchronize()
    got = q.cpu().numpy()
    assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max()
    assert np.array_equal(got, want), "generated source-term kernel differs from the oracle in the last bits"
    np.testing.assert_allclose(lam.cpu().numpy(), lam_o, rtol=1e-12)
