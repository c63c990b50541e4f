"""CPU-only checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/exahype_cuda.h declares, validates arguments like KernelBuilder.viable(), and fails loudly
(no CPU fallback) when asked to compute without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

from exahype_b200 import runtime

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    return runtime.load()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "exahype_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(exahype_cuda_\w+)\s*\(", text)))


def test_header_symbols_are_all_exported(lib):
    names = declared_symbols()
    assert len(names) >= 18
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/exahype_cuda.h but not exported"


def test_version_and_registry(lib):
    assert lib.exahype_cuda_version() == 2
    inst = runtime.committed_instantiations()
    keys = {(i["model"], i["dim"], i["patch_size"], i["n_real"], i["n_aux"], i["dtype"]) for i in inst}
    # BASELINE.json configs C1..C4 and the shape of the reference's committed kernel
    for want in [("euler", 2, 3, 4, 0, "f64"), ("euler", 2, 16, 4, 0, "f64"), ("euler", 3, 8, 5, 0, "f64"),
                 ("swe", 2, 32, 3, 1, "f64"), ("swe", 2, 32, 3, 1, "f32"), ("euler", 2, 4, 5, 5, "f64")]:
        assert want in keys


def test_argument_validation_mirrors_viable(lib):
    def rc(**kw):
        base = dict(model=0, dtype=0, dim=3, patch_size=8, halo=1, n_real=5, n_aux=0, flags=0)
        base.update(kw)
        c = runtime.FvConfig(*[base[k] for k in ("model", "dtype", "dim", "patch_size", "halo", "n_real", "n_aux", "flags")])
        return lib.exahype_cuda_fv_step(ctypes.byref(c), None, None, 0, 0.0, None, None, None)
    assert rc(dim=4) == -1 and b"viability" in lib.exahype_cuda_last_error()
    assert rc(patch_size=0) == -1
    assert rc(halo=-1) == -1
    assert rc(halo=0) == -1
    assert rc(flags=1 << 9) == -1
    # the unknowns-only output is a form of the un-haloed one
    assert rc(flags=runtime.FLAG_OUTPUT_UNKNOWNS_ONLY) == -1 and b"UNHALOED" in lib.exahype_cuda_last_error()
    assert rc(flags=runtime.FLAG_OUTPUT_UNKNOWNS_ONLY | runtime.FLAG_OUTPUT_UNHALOED) == 0
    assert rc(patch_size=7) == -2 and b"CUDAPrinter" in lib.exahype_cuda_last_error()
    assert rc() == 0                      # zero patches: nothing to do, no device needed
    with pytest.raises(Exception, match="viability"):
        runtime.PatchUpdate(dim=5)


def test_no_cpu_fallback_without_a_device(lib):
    if runtime.device_count() > 0:
        pytest.skip("a CUDA device is present")
    upd = runtime.PatchUpdate(model="euler", dim=2, patch_size=3, n_real=4)
    Q = np.ones(upd.in_shape(2))
    with pytest.raises(runtime.ExaHyPECudaError) as e:
        upd.time_step(Q, 0.1)
    assert e.value.code == -3
    assert np.all(Q == 1.0)               # nothing was computed on the host


def test_algorithmic_bytes_match_baseline_table():
    # BASELINE.md section 3
    assert runtime.PatchUpdate("euler", 2, 3, 1, 4, 0).algorithmic_bytes_per_patch == 968
    assert runtime.PatchUpdate("euler", 2, 16, 1, 4, 0).algorithmic_bytes_per_patch == 18440
    assert runtime.PatchUpdate("euler", 3, 8, 1, 5, 0).algorithmic_bytes_per_patch == 56328
    assert runtime.PatchUpdate("swe", 2, 32, 1, 3, 1).algorithmic_bytes_per_patch == 61448
    assert runtime.PatchUpdate("swe", 2, 32, 1, 3, 1, dtype="f32").algorithmic_bytes_per_patch == 30724
