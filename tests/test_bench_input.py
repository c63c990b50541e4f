"""bench.py generates its input on the device with torch integer ops; it must be the oracle's synthetic state bit for bit
(so the benchmark runs on exactly the data the parity tests cover).  Checked here on CPU tensors."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from exahype_b200 import runtime  # noqa: E402


@pytest.mark.parametrize("wl", ["c3", "c2", "c4", "c4f32", "c1"])
def test_device_generator_equals_oracle_fill(oracle, wl):
    model, dim, P, h, nr, na, dtype, _, _ = bench.WORKLOADS[wl]
    upd = runtime.PatchUpdate(model, dim, P, h, nr, na, dtype=dtype)
    cfg = oracle.OracleConfig(dim=dim, patch_size=P, halo=h, n_real=nr, n_aux=na,
                              model=oracle.MODEL_EULER if model == "euler" else oracle.MODEL_SWE)
    tdt = torch.float64 if dtype == "f64" else torch.float32
    npdt = np.float64 if dtype == "f64" else np.float32
    got = bench.synthetic_on_device(torch, upd, 5, 9, tdt, device="cpu").numpy()
    want = oracle.fill_synthetic(cfg, 9, dtype=npdt, first_patch=5)
    assert got.shape == want.shape and np.array_equal(got, want)
