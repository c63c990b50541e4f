"""TEST INFRASTRUCTURE -- NOT PART OF THE PRODUCT PATH.

ctypes loader for the CPU oracle (``oracle/fv_rusanov_oracle.c``) and, when it was built in the
dev container, the reference's own compiled kernel (``oracle/_ref/libexahype_ref.so``).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this package.  ``exahype_b200`` never does.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

MODEL_EULER, MODEL_SWE, MODEL_SWE_SOURCE = 0, 1, 2
RANGES_HEAD, RANGES_COMMITTED = 0, 1
DISS_VAR0, DISS_ALL = 0, 1

SEED = 20240601


class _Cfg(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in
                ("dim", "patch_size", "halo", "n_real", "n_aux", "model", "ranges", "diss")]


@dataclass(frozen=True)
class OracleConfig:
    dim: int
    patch_size: int
    halo: int = 1
    n_real: int = 4
    n_aux: int = 0
    model: int = MODEL_EULER
    ranges: int = RANGES_HEAD
    diss: int = DISS_VAR0

    @property
    def side(self) -> int:
        return self.patch_size + 2 * self.halo

    @property
    def n_var(self) -> int:
        return self.n_real + self.n_aux

    @property
    def cells_per_patch(self) -> int:
        return self.side ** self.dim

    @property
    def values_per_patch(self) -> int:
        return self.cells_per_patch * self.n_var

    def shape(self, n_patches: int):
        return (n_patches,) + (self.side,) * self.dim + (self.n_var,)

    def _c(self) -> _Cfg:
        return _Cfg(self.dim, self.patch_size, self.halo, self.n_real, self.n_aux,
                    self.model, self.ranges, self.diss)


def build(fast: bool = True) -> None:
    """Compile the oracle (and oracle/_ref when /root/reference is present)."""
    subprocess.run(["make", "-C", _HERE, "-s", "all"], check=True,
                   stdout=subprocess.DEVNULL)


_libs: dict = {}


def _lib(fast: bool = False) -> ctypes.CDLL:
    name = "libfv_oracle_fast.so" if fast else "libfv_oracle.so"
    if name not in _libs:
        path = os.path.join(_HERE, name)
        if not os.path.exists(path):
            build()
        lib = ctypes.CDLL(path)
        for sfx, ct in (("f64", ctypes.c_double), ("f32", ctypes.c_float)):
            f = getattr(lib, f"fvo_step_{sfx}")
            f.restype = ctypes.c_int
            f.argtypes = [ctypes.POINTER(_Cfg), ctypes.c_void_p, ctypes.c_int64, ct,
                          ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
            g = getattr(lib, f"fvo_fill_sin_{sfx}")
            g.restype = None
            g.argtypes = [ctypes.c_void_p, ctypes.c_int64]
            s = getattr(lib, f"fvo_fill_synthetic_{sfx}")
            s.restype = None
            s.argtypes = [ctypes.POINTER(_Cfg), ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64,
                          ctypes.c_uint64]
        lib.fvo_fnv1a64_words.restype = ctypes.c_uint64
        lib.fvo_fnv1a64_words.argtypes = [ctypes.c_void_p, ctypes.c_int64]
        lib.fvo_max_threads.restype = ctypes.c_int
        _libs[name] = lib
    return _libs[name]


def _sfx(dtype) -> str:
    dtype = np.dtype(dtype)
    if dtype == np.float64:
        return "f64"
    if dtype == np.float32:
        return "f32"
    raise TypeError(f"oracle supports float64/float32, got {dtype}")


def max_threads() -> int:
    return int(_lib().fvo_max_threads())


def step(cfg: OracleConfig, Q: np.ndarray, dt: float, nthreads: int = 1, fast: bool = False):
    """In-place patch update of ``Q`` (C-contiguous, AoS).  Returns (lambda_patch, lambda_max)."""
    if not Q.flags["C_CONTIGUOUS"]:
        raise ValueError("Q must be C-contiguous")
    sfx = _sfx(Q.dtype)
    if Q.size % cfg.values_per_patch:
        raise ValueError("Q size is not a whole number of patches")
    n_patches = Q.size // cfg.values_per_patch
    lam = np.zeros(n_patches, dtype=Q.dtype)
    lmax = np.zeros(1, dtype=Q.dtype)
    c = cfg._c()
    rc = getattr(_lib(fast), f"fvo_step_{sfx}")(ctypes.byref(c), Q.ctypes.data, n_patches, float(dt),
                                                 lam.ctypes.data, lmax.ctypes.data, int(nthreads))
    if rc:
        raise ValueError(f"oracle rejected the configuration (rc={rc})")
    return lam, lmax[0]


def fill_sin(cfg: OracleConfig, n_patches: int, dtype=np.float64) -> np.ndarray:
    """``Q[i] = sin(3.141*i/N)`` over the flat batch (correctness_test.cpp:102-106)."""
    Q = np.empty(cfg.shape(n_patches), dtype=dtype)
    getattr(_lib(), f"fvo_fill_sin_{_sfx(dtype)}")(Q.ctypes.data, Q.size)
    return Q


def fill_synthetic(cfg: OracleConfig, n_patches: int, dtype=np.float64, first_patch: int = 0,
                   seed: int = SEED) -> np.ndarray:
    """Counter-based admissible state of SURVEY.md section 8d for patches [first_patch, +n_patches)."""
    Q = np.empty(cfg.shape(n_patches), dtype=dtype)
    c = cfg._c()
    getattr(_lib(), f"fvo_fill_synthetic_{_sfx(dtype)}")(
        ctypes.byref(c), Q.ctypes.data, first_patch * cfg.cells_per_patch,
        n_patches * cfg.cells_per_patch, seed)
    return Q


def fnv1a64(a: np.ndarray) -> str:
    """FNV-1a-64 over the 8-byte words of ``a`` (hex)."""
    a = np.ascontiguousarray(a)
    if a.nbytes % 8:
        raise ValueError("hash is defined over 8-byte words")
    return format(_lib().fvo_fnv1a64_words(a.ctypes.data, a.nbytes // 8), "016x")


# ---------------------------------------------------------------------------------------------
# the reference's own compiled kernel (dev container only; prebuilt .so travels to the GPU box)

def reference_available() -> bool:
    return os.path.exists(os.path.join(_HERE, "_ref", "libexahype_ref.so"))


def _ref() -> ctypes.CDLL:
    if "ref" not in _libs:
        lib = ctypes.CDLL(os.path.join(_HERE, "_ref", "libexahype_ref.so"))
        lib.ref_time_step.restype = None
        lib.ref_time_step.argtypes = [ctypes.c_void_p, ctypes.c_double]
        lib.ref_flux.restype = None
        lib.ref_flux.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        lib.ref_max_eigenvalue.restype = ctypes.c_double
        lib.ref_max_eigenvalue.argtypes = [ctypes.c_void_p, ctypes.c_int]
        _libs["ref"] = lib
    return _libs["ref"]


#: configuration hard-wired into the reference's committed kernel (Unit test/test.cpp:4-8)
REFERENCE_CONFIG = OracleConfig(dim=2, patch_size=4, halo=1, n_real=5, n_aux=5,
                                ranges=RANGES_COMMITTED, diss=DISS_VAR0)


def reference_time_step(Q: np.ndarray, dt: float) -> None:
    """``time_step(Q, dt)`` of the reference's committed generated kernel, in place (360 doubles)."""
    if Q.dtype != np.float64 or Q.size != 360 or not Q.flags["C_CONTIGUOUS"]:
        raise ValueError("the committed reference kernel takes exactly 360 contiguous doubles")
    _ref().ref_time_step(Q.ctypes.data, float(dt))


def reference_flux(q: np.ndarray, normal: int, n_out: int = 5) -> np.ndarray:
    q = np.ascontiguousarray(q, dtype=np.float64)
    F = np.zeros(n_out)
    _ref().ref_flux(q.ctypes.data, int(normal), F.ctypes.data)
    return F


def reference_max_eigenvalue(q: np.ndarray, normal: int) -> float:
    q = np.ascontiguousarray(q, dtype=np.float64)
    return float(_ref().ref_max_eigenvalue(q.ctypes.data, int(normal)))


def reference_compiled_rate(min_seconds: float = 2.0):
    """Interior cell-updates/s of the reference's own committed kernel (``Unit test/test.cpp`` + ``Functions.cpp``,
    compiled from where they lie: ``oracle/_ref/libexahype_ref_fast.so``, ``-O3 -march=native``) on its hard-wired shape --
    one 4x4 2-D patch with 5 + 5 variables, one thread, as the reference runs.  Returns ``(rate, description)`` or
    ``None`` when the prebuilt library is not there."""
    import time
    path = os.path.join(_HERE, "_ref", "libexahype_ref_fast.so")
    if not os.path.exists(path):
        return None
    lib = ctypes.CDLL(path)
    lib.ref_time_step_repeat.restype = None
    lib.ref_time_step_repeat.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_double, ctypes.c_long]
    q0 = fill_sin(REFERENCE_CONFIG, 1).ravel().copy()
    q = q0.copy()
    reps, calls, elapsed = 20000, 0, 0.0
    lib.ref_time_step_repeat(q0.ctypes.data, q.ctypes.data, 1.0, 1000)
    while elapsed < min_seconds:
        t0 = time.perf_counter()
        lib.ref_time_step_repeat(q0.ctypes.data, q.ctypes.data, 1.0, reps)
        elapsed += time.perf_counter() - t0
        calls += reps
    return 16 * calls / elapsed, (f"reference's committed kernel (Unit test/test.cpp + Functions.cpp, g++ -O3 -march=native), "
                                  f"its hard-wired 4x4 2-D patch with 5+5 variables, {calls} calls in {elapsed:.1f} s, 1 thread, "
                                  f"incl. its five new[]/delete[] per call")
