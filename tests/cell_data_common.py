"""Shared by the CellData-boundary tests.

``declare()`` states, for any shape, the declaration that the reference's user script ``examples/kernel-generator.py:6-45``
makes for 2-D 4x4 patches: an ExaHyPE2 ``CellData`` object with members QIn / QOut / dt / t / cellCentre / cellSize and solver
functions with the signature ``flux(Q, x, h, t, dt, normal, F)``.  tests/test_cell_data_cpu.py runs the reference's script
itself (from /root/reference, where that exists) and checks that both produce the same statements and the same code, so
the GPU tests -- which cannot read /root/reference -- exercise the reference's declaration.  ``compile_generated_cpp()``
builds what ``CPPPrinter`` emits for it with g++ against the minimal fake of ExaHyPE2's types in
``tests/cpp/fake_exahype2.h``: the CPU side of the parity tests.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

# The solver of tests/cpp/fake_exahype2.h as device source: same formulas, same evaluation order.
def device_solver(dim: int) -> str:
    return _DEVICE_SOLVER.replace("DIMENSIONS", str(dim))


_DEVICE_SOLVER = """
template <class T> __device__ void flux(const T* Q, const T* x, const T* h, T t, T dt, int normal, T* F) {
  constexpr int D = DIMENSIONS;
  const T irho = T(1.0) / Q[0];
  T ke = Q[1] * Q[1] + Q[2] * Q[2];
  if (D == 3) ke = ke + Q[3] * Q[3];
  const T p = (T(1.4) - 1) * (Q[D + 1] - T(0.5) * irho * ke);
  const T coeff = irho * Q[normal + 1];
  const T w = T(1.0) + T(0.01) * x[normal] + T(0.1) * h[0] + T(0.001) * t + T(0.5) * dt;
  for (int v = 0; v <= D; ++v) F[v] = coeff * Q[v] * w;
  F[D + 1] = (coeff * Q[D + 1] + coeff * p) * w;
  F[normal + 1] += p;
}
template <class T> __device__ T flux(const T* Q, const T* x, const T* h, T t, T dt, int normal) {
  const T irho = T(1.0) / Q[0];
  return ::exahype::fv_abs(Q[normal + 1] * irho) + T(0.01) * x[0] + h[1] + t + dt;
}
"""


SOLVER = "benchmarks::exahype2::kernelbenchmarks::repositories::instanceOfFVRusanovSolver"


def declare(dim=2, patch_size=4, halo_size=1, n_real=4, n_aux=0):
    """The CellData-boundary kernel for one shape (same objects, parents and statements as kernel-generator.py:6-45)."""
    from exahype import KernelBuilder
    k = KernelBuilder(dim=dim, patch_size=patch_size, halo_size=halo_size, n_real=n_real, n_aux=n_aux)
    data = k.item('patchData', in_type='::exahype2::CellData&')
    k.const('timingComputeKernel', in_type='::tarch::timing::Measurement&')
    q = k.item('QOut', parent=data)
    qc = k.item('QIn', parent=data)
    f = k.directional_item('tmp_flx')
    lam = k.directional_item('tmp_eigen', struct=False)
    dt, t = k.const('dt', parent=data), k.const('t', parent=data)
    normal = k.directional_const('normal', tuple(range(dim)))
    centre, size = k.const('cellCentre', parent=data), k.const('cellSize', parent=data)
    flux = k.function('flux', parent=SOLVER)
    k.function('maxEigenvalue', parent=SOLVER)
    mx = k.function('max')
    vol_centre = k.function('getVolumeCentre', parent='exahype2::fv::')
    vol_size = k.function('getVolumeSize', parent='exahype2::fv::')
    P = k.all_items["patch_size"]
    index = {k.all_items[a] for a in ("i", "j", "k")[:dim]}
    k.single(qc[0], q[0])
    k.directional(flux(qc[0], vol_centre(centre, size, P, index), vol_size(size, P), t, dt, normal, f[0]))
    k.directional(lam[0], flux(qc[0], vol_centre(centre, size, P), vol_size(size, P), t, dt, normal))
    k.directional(qc[0], qc[0] + 0.5 * (f[-1] - f[1]))
    left = -mx(lam[-1], lam[0]) * (q[0] - q[-1])
    right = -mx(lam[1], lam[0]) * (q[0] - q[1])
    k.directional(qc[0], qc[0] + 0.5 * dt * (left - right), struct=True)
    k.single(q[0], qc[0])
    return k


def compile_generated_cpp(tmp_path, dim: int, generated="generated_kernel.cpp"):
    """g++ the harness around the generated unit ``tmp_path/generated``; returns ``run(q, centre, size, t, dt)``, which
    updates the haloed batch ``q`` in place."""
    lib = os.path.join(tmp_path, "libcell_data_cpu.so")
    cmd = ["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-shared", "-fPIC", f"-DDimensions={dim}",
           f'-DGENERATED_KERNEL="{generated}"', "-I", os.path.join(HERE, "cpp"), "-I", str(tmp_path),
           os.path.join(HERE, "cpp", "cell_data_harness.cpp"), "-o", lib]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, "the generated C++ does not compile:\n" + r.stderr[-4000:]
    fn = ctypes.CDLL(lib).run_cell_data
    vp = ctypes.c_void_p
    fn.argtypes = [ctypes.c_int, ctypes.c_longlong, vp, vp, vp, vp, vp, vp]

    def run(q, centre, size, t, dt):
        n = q.shape[0]
        per = int(np.prod(q.shape[1:]))
        scratch = np.zeros(per)
        for a in (q, centre, size, t, dt):
            assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
        fn(n, per, q.ctypes.data, scratch.ctypes.data, centre.ctypes.data, size.ctypes.data, t.ctypes.data, dt.ctypes.data)
        return q
    return run


def patch_geometry(n, dim, seed=3):
    rng = np.random.default_rng(seed)
    centre = rng.uniform(-1.0, 1.0, (n, dim))
    size = rng.uniform(0.05, 0.2, (n, dim))
    t = rng.uniform(0.0, 2.0, n)
    dt = rng.uniform(0.001, 0.02, n)
    return centre, size, t, dt
