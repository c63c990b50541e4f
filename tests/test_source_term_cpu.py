"""Source term on the CPU: the oracle's definition against an independent numpy statement, the
C++ that CPPPrinter generates from the declaration (with the user's sourceTerm in Functions.h style) against the oracle,
bit for bit, and the CUDA back-end's recognition of the two statements."""
import ctypes
import subprocess

import numpy as np
import pytest

import source_term_common as S


def cfg(oracle, P, model=None):
    return oracle.OracleConfig(dim=2, patch_size=P, halo=1, n_real=3, n_aux=3,
                               model=oracle.MODEL_SWE_SOURCE if model is None else model)


@pytest.mark.parametrize("npdt", [np.float64, np.float32])
def test_oracle_source_model_is_the_source_free_step_plus_dt_times_s(oracle, npdt):
    q = oracle.fill_synthetic(cfg(oracle, 8), 6, dtype=npdt)
    with_source, without = q.copy(), q.copy()
    lam_a, lmax_a = oracle.step(cfg(oracle, 8), with_source, 0.01)
    lam_b, lmax_b = oracle.step(cfg(oracle, 8, oracle.MODEL_SWE), without, 0.01)
    dt = npdt(0.01)
    s = np.zeros_like(q[..., :3])
    gh = npdt(9.81) * q[..., 0]
    s[..., 1] = -gh * q[..., 4]
    s[..., 2] = -gh * q[..., 5]
    want = without.copy()
    inner = (slice(None), slice(1, -1), slice(1, -1))
    want[inner + (slice(0, 3),)] = without[inner + (slice(0, 3),)] + dt * s[inner]
    assert np.array_equal(with_source, want)
    assert np.array_equal(lam_a, lam_b) and lmax_a == lmax_b          # the eigenvalue is that of the input state
    assert not np.array_equal(with_source, without)
    assert np.array_equal(with_source[..., 3:], q[..., 3:])           # aux pass through


def test_generated_cpp_with_source_statements_equals_the_oracle(tmp_path, oracle):
    from exahype.printers import CPPPrinter
    k = S.declare(patch_size=8)
    (tmp_path / "Functions.h").write_text(S.HOST_HEADER)
    (tmp_path / "Functions.cpp").write_text(S.HOST_FUNCTIONS)
    CPPPrinter(k).file(file_name=str(tmp_path / "time_step.cpp"), header_file_name="Functions.h")
    lib = str(tmp_path / "libsrc.so")
    r = subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-I", str(tmp_path),
                        str(tmp_path / "time_step.cpp"), str(tmp_path / "Functions.cpp"), "-o", lib], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    fn = ctypes.CDLL(lib)._Z9time_stepPdd                 # void time_step(double*, double)
    fn.argtypes = [ctypes.c_void_p, ctypes.c_double]
    q0 = oracle.fill_synthetic(cfg(oracle, 8), 5)
    want = q0.copy()
    oracle.step(cfg(oracle, 8), want, 0.01)
    got = q0.copy()
    for p in range(got.shape[0]):                         # the declaration has n_patches = 1
        fn(got[p].ctypes.data, 0.01)
    assert np.array_equal(got, want)


def test_cuda_printer_recognises_the_source_statements(tmp_path):
    from exahype.printers import CUDAPrinter
    from exahype_b200.printers.CUDAPrinter import UnsupportedKernel
    cu = CUDAPrinter(S.declare(patch_size=32), model="swe_source")
    p = cu.program
    assert (p.source_fn, p.source_tmp, p.source_update) == ("sourceTerm", "tmp_source", "dt*s + qc")
    assert p.roles[-3:] == ["source call", "source update", "copy-out"]
    assert "::exahype::SweSourcePhysics<3, 3>" in cu.code and cu.template == "march"
    user = CUDAPrinter(S.declare(patch_size=8, device_source=True), function_name="swe_source_user")
    assert "static constexpr bool HAS_SOURCE = true;" in user.code and "user::sourceTerm(q, S);" in user.code
    assert "using Update = ::exahype::RusanovUpdate;" in user.code      # the declaration's update statements are the reference's
    assert user.build(directory=str(tmp_path)).lib_path                      # cross-compiles for sm_100a
    with pytest.raises(UnsupportedKernel):
        CUDAPrinter(S.declare(patch_size=8), model="swe")                    # source statements need the source family
    with pytest.raises(UnsupportedKernel):
        CUDAPrinter(S.declare(patch_size=8, source=False), model="swe_source")
