"""The ExaHyPE2 CellData boundary on the GPU: the kernel that CUDAPrinter generates from the
reference's examples/kernel-generator.py declaration -- CellData members QIn / QOut / dt / t / cellCentre / cellSize, solver
functions flux(Q, x, h, t, dt, normal, F) -- against the C++ that CPPPrinter generates from the SAME declaration, compiled
with g++ over a minimal fake of ExaHyPE2's types (tests/cpp/fake_exahype2.h; pinned to the oracle by
tests/test_cell_data_cpu.py).  The test solver depends on every argument of the signature, so a context value that does
not reach the functor -- or reaches it for the wrong cell -- fails the comparison.

Bound: 1e-12 relative (BASELINE.json north_star); asserted bitwise, which the contraction-free build delivers.
"""
import numpy as np
import pytest

import cell_data_common as C

pytestmark = pytest.mark.gpu
RTOL = 1e-12


@pytest.fixture(scope="module")
def torch():
    import torch
    assert torch.cuda.is_available()
    return torch


def _cpu_reference(tmp_path, kernel, dim, q0, centre, size, t, dt):
    from exahype.printers import CPPPrinter
    CPPPrinter(kernel).file(file_name=str(tmp_path / "generated_kernel.cpp"))
    run = C.compile_generated_cpp(str(tmp_path), dim)
    return run(q0.copy(), centre, size, t, dt)


@pytest.mark.parametrize("dim,P,n", [(2, 4, 37), (3, 4, 19), (2, 16, 11)])
@pytest.mark.parametrize("unhaloed", [False, True])
def test_generated_cell_data_kernel_equals_generated_cpp(torch, oracle, tmp_path, dim, P, n, unhaloed):
    from exahype.printers import CUDAPrinter
    nr = dim + 2
    kernel = C.declare(dim=dim, patch_size=P, n_real=nr)
    cfg = oracle.OracleConfig(dim=dim, patch_size=P, halo=1, n_real=nr, n_aux=0)
    q0 = oracle.fill_synthetic(cfg, n)
    centre, size, t, dt = C.patch_geometry(n, dim)
    want = _cpu_reference(tmp_path, kernel, dim, q0, centre, size, t, dt)
    assert not np.array_equal(want, q0)

    kernel.all_items["flux"].deviceBody(C.device_solver(dim))
    cu = CUDAPrinter(kernel)
    assert cu.context and cu.template == "cell"
    gk = cu.build()

    # patches scattered through a pool (CellData::QIn[p] are independent pointers), with slack that must stay untouched
    per = q0[0].size
    stride = per + 6                                          # 8-byte slack keeps every patch 16-byte aligned
    perm = np.random.default_rng(1).permutation(n)
    pool = torch.full((n * stride,), -7.0, dtype=torch.float64, device="cuda")
    for p in range(n):
        pool[perm[p] * stride: perm[p] * stride + per] = torch.from_numpy(q0[p].ravel()).cuda()
    in_ptrs = torch.tensor([pool.data_ptr() + int(perm[p]) * stride * 8 for p in range(n)], dtype=torch.int64, device="cuda")
    if unhaloed:
        out = torch.full((n,) + (P,) * dim + (nr,), -3.0, dtype=torch.float64, device="cuda")
        out_ptrs = torch.tensor([out.data_ptr() + p * out[0].numel() * 8 for p in range(n)], dtype=torch.int64, device="cuda")
    else:
        out_ptrs = in_ptrs                                    # in place, like the declaration's QOut
    lam_patch = torch.zeros(n, dtype=torch.float64, device="cuda")
    lam_max = torch.zeros(1, dtype=torch.float64, device="cuda")
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    gk.step_cell_data(in_ptrs, out_ptrs, dt_patch=dev(dt), max_eigenvalue=lam_patch, lambda_max=lam_max,
                      cell_centre=dev(centre), cell_size=dev(size), t_patch=dev(t), unhaloed=unhaloed)
    torch.cuda.synchronize()
    interior = (slice(None),) + (slice(1, -1),) * dim + (slice(None),)
    if unhaloed:
        got = out.cpu().numpy()
        ref = want[interior]
    else:
        host = pool.cpu().numpy().reshape(n, stride)
        assert np.all(host[:, per:] == -7.0), "slack between the gathered patches was written"
        got = np.stack([host[perm[p], :per].reshape(q0[0].shape) for p in range(n)])
        ref = want
    scale = np.abs(ref).max()
    assert np.abs(got - ref).max() <= RTOL * scale
    assert np.array_equal(got, ref), "generated CUDA differs from the generated C++ in the last bits"
    assert float(lam_max.item()) == float(lam_patch.max().item()) > 0.0


def test_generated_step_validates_like_patch_update(torch, oracle):
    """the generated binding checks dtype / sizes / aliasing like PatchUpdate.step instead of reading out of bounds."""
    from exahype.printers import CUDAPrinter
    kernel = C.declare(dim=2, patch_size=4, n_real=4)
    gk = CUDAPrinter(kernel, model="euler").build()
    q = torch.ones(gk.in_shape(3), dtype=torch.float64, device="cuda")
    with pytest.raises(ValueError):
        gk.step(q.float(), None, 0.1)                                       # fp32 tensor into an fp64 kernel
    with pytest.raises(ValueError):
        gk.step(q, torch.empty(5, dtype=torch.float64, device="cuda"), 0.1, unhaloed=True)   # q_out too small
    with pytest.raises(ValueError):
        gk.step(q, None, 0.1, unhaloed=True)                                # un-haloed output needs its own buffer
    with pytest.raises(ValueError):
        gk.step(q, None, 0.1, lambda_patch=torch.empty(1, dtype=torch.float64, device="cuda"))
    gk.step(q, None, 0.1)
    torch.cuda.synchronize()
