"""Host binding: ``ctypes`` over ``libexahype_cuda.so`` (C ABI in ``include/exahype_cuda.h``).

The reference's host-side call is ``time_step(Q, dt)`` on one patch
(``Unit test/test.h:3``, called from ``Unit test/correctness_test.cpp:195``).  :class:`PatchUpdate` keeps that
call shape -- ``time_step(Q, dt)`` on a host array updates it in place -- and adds the device-resident form
``step(q_in, q_out, dt)`` on CUDA tensors for whole batches.

PyTorch is used for device memory and streams only.  There is no CPU fallback: if the library is missing it is
built with nvcc, and if that is impossible, or no CUDA device is present at call time, an error is raised.
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# EXAHYPE_CUDA_LIB points the binding at another build of the same ABI (kernel-tuning experiments)
LIB_PATH = os.environ.get("EXAHYPE_CUDA_LIB") or os.path.join(_HERE, "libexahype_cuda.so")

MODEL = {"euler": 0, "swe": 1, "swe_source": 2}   # swe_source: + bathymetry source term, aux = (b, db/dx, db/dy)
DTYPE = {"f64": 0, "f32": 1}
FLAG_DISSIPATION_ALL = 1 << 0
FLAG_OUTPUT_UNHALOED = 1 << 1
FLAG_LAMBDA_ACCUMULATE = 1 << 2
FLAG_KERNEL_CELL = 1 << 3
FLAG_FAST_ARITHMETIC = 1 << 4
FLAG_OUTPUT_UNKNOWNS_ONLY = 1 << 5

ERR_NAMES = {-1: "INVALID_ARGUMENT", -2: "NO_INSTANTIATION", -3: "CUDA", -4: "NCCL", -5: "UNAVAILABLE", -6: "TIMEOUT"}


class ExaHyPECudaError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libexahype_cuda: {ERR_NAMES.get(code, code)}: {message}")
        self.code = code


class FvConfig(ctypes.Structure):
    """``exahype_fv_config`` of include/exahype_cuda.h."""
    _fields_ = [("model", ctypes.c_int32), ("dtype", ctypes.c_int32), ("dim", ctypes.c_int32),
                ("patch_size", ctypes.c_int32), ("halo", ctypes.c_int32), ("n_real", ctypes.c_int32),
                ("n_aux", ctypes.c_int32), ("flags", ctypes.c_uint32)]


class CellData(ctypes.Structure):
    """``exahype_cell_data`` of include/exahype_cuda.h (ExaHyPE2's ``CellData``: per-patch pointers and time steps)."""
    _fields_ = [("n_patches", ctypes.c_int64), ("q_in", ctypes.c_void_p), ("q_out", ctypes.c_void_p),
                ("dt", ctypes.c_void_p), ("max_eigenvalue", ctypes.c_void_p), ("cell_centre", ctypes.c_void_p),
                ("cell_size", ctypes.c_void_p), ("t", ctypes.c_void_p)]


_lib: Optional[ctypes.CDLL] = None


def load(path: Optional[str] = None) -> ctypes.CDLL:
    """Load (building first if necessary) ``libexahype_cuda.so`` and declare the ABI."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        from . import build as _build   # nvcc cross-compiles; raises when nvcc is missing
        _build.build()
    lib = ctypes.CDLL(path)
    c_cfg = ctypes.POINTER(FvConfig)
    vp, i64, i32, dbl = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_double
    lib.exahype_cuda_version.restype = i32
    lib.exahype_cuda_last_error.restype = ctypes.c_char_p
    lib.exahype_cuda_device_count.restype = i32
    lib.exahype_cuda_fv_supported.argtypes = [c_cfg]
    lib.exahype_cuda_fv_list.argtypes = [c_cfg, i32]
    lib.exahype_cuda_fv_step.argtypes = [c_cfg, vp, vp, i64, dbl, vp, vp, vp]
    lib.exahype_cuda_fv_step_allreduce.argtypes = [c_cfg, vp, vp, vp, i64, dbl, vp, vp, vp]
    lib.exahype_cuda_fv_step_cell_data.argtypes = [c_cfg, ctypes.POINTER(CellData), dbl, vp, vp]
    lib.exahype_cuda_time_step_host.argtypes = [c_cfg, vp, vp, i64, dbl, vp, vp]
    lib.exahype_cuda_host_pipeline_configure.argtypes = [i64, i32]
    lib.exahype_cuda_launch_count.restype = i64
    lib.exahype_cuda_fill_synthetic.argtypes = [c_cfg, vp, i64, i64, ctypes.c_uint64, vp]
    ip = ctypes.POINTER(ctypes.c_int)
    lib.exahype_cuda_fv_launch_info.argtypes = [c_cfg, i64, ip, ip, ip, ip]
    lib.exahype_cuda_nccl_unique_id.argtypes = [vp]
    lib.exahype_cuda_comm_init.argtypes = [ctypes.POINTER(vp), vp, i32, i32]
    lib.exahype_cuda_comm_destroy.argtypes = [vp]
    lib.exahype_cuda_allreduce_max.argtypes = [vp, vp, i64, i32, vp]
    lib.exahype_cuda_peer_reducer_create.argtypes = [ctypes.POINTER(vp), i32, i32]
    lib.exahype_cuda_peer_reducer_local_handle.argtypes = [vp, vp]
    lib.exahype_cuda_peer_reducer_connect.argtypes = [vp, vp]
    lib.exahype_cuda_peer_reducer_allreduce_max.argtypes = [vp, vp, i32, vp]
    lib.exahype_cuda_peer_reducer_status.argtypes = [vp, ip]
    lib.exahype_cuda_peer_reducer_destroy.argtypes = [vp]
    lib.exahype_cuda_peer_reducer_set_timeout.argtypes = [vp, dbl]
    lib.exahype_cuda_peer_reducer_connect_local.argtypes = [ctypes.POINTER(vp), i32]
    lib.exahype_cuda_peer_reducer_enable_trace.argtypes = [vp, i32]
    lib.exahype_cuda_peer_reducer_read_trace.argtypes = [vp, ctypes.c_uint64, i32, vp]
    lib.exahype_cuda_time_loop_create.argtypes = [ctypes.POINTER(vp), i32, vp, dbl, dbl, i64]
    lib.exahype_cuda_fv_step_time_loop.argtypes = [c_cfg, vp, vp, vp, i64, vp, vp]
    lib.exahype_cuda_time_loop_flush.argtypes = [vp, vp]
    lib.exahype_cuda_time_loop_history.argtypes = [vp, i64, i64, vp]
    lib.exahype_cuda_time_loop_steps.argtypes = [vp]
    lib.exahype_cuda_time_loop_steps.restype = i64
    lib.exahype_cuda_time_loop_dt_device.argtypes = [vp, ctypes.POINTER(vp)]
    lib.exahype_cuda_time_loop_destroy.argtypes = [vp]
    for t, ct in (("f64", ctypes.c_double), ("f32", ctypes.c_float)):
        for name in (f"exahype_cuda_fv_step_euler_2d_{t}", f"exahype_cuda_fv_step_euler_3d_{t}",
                     f"exahype_cuda_fv_step_swe_2d_{t}"):
            if hasattr(lib, name):
                getattr(lib, name).argtypes = [vp, vp, i64, i32, i32, i32, ct, vp, vp, ctypes.c_uint, vp]
    if path == LIB_PATH:
        _lib = lib
    return lib


def check(rc: int, lib: Optional[ctypes.CDLL] = None) -> None:
    if rc != 0:
        lib = lib or load()
        raise ExaHyPECudaError(rc, lib.exahype_cuda_last_error().decode())


def device_count() -> int:
    return int(load().exahype_cuda_device_count())


def launch_count() -> int:
    return int(load().exahype_cuda_launch_count())


def committed_instantiations():
    lib = load()
    n = lib.exahype_cuda_fv_list(None, 0)
    arr = (FvConfig * n)()
    lib.exahype_cuda_fv_list(arr, n)
    inv_m = {v: k for k, v in MODEL.items()}
    inv_t = {v: k for k, v in DTYPE.items()}
    return [dict(model=inv_m[c.model], dtype=inv_t[c.dtype], dim=c.dim, patch_size=c.patch_size, halo_size=c.halo,
                 n_real=c.n_real, n_aux=c.n_aux) for c in arr]


def _np_dtype(dtype: str):
    return np.float64 if dtype == "f64" else np.float32


@dataclass
class PatchUpdate:
    """One configured batched patch update (``time_step`` of the reference, for a batch, on the GPU).

    ``dissipation='var0'`` is the reference-emitted behaviour (``Unit test/test.cpp:81,90``), ``'all'`` what the
    declaration intends.  ``output='haloed'`` writes interior cells of a buffer shaped like the input (in place when
    ``q_out is q_in``); ``'unhaloed'`` writes ``[n_patches, P.., n_var]`` (ExaHyPE2's ``QOut``); ``'unknowns'`` writes
    ``[n_patches, P.., n_real]`` -- the un-haloed form without the auxiliary variables, which a step never changes
    (``EXAHYPE_FLAG_OUTPUT_UNKNOWNS_ONLY``: 13 % less DRAM traffic on 32x32 shallow-water patches).
    """
    model: str = "euler"
    dim: int = 3
    patch_size: int = 8
    halo_size: int = 1
    n_real: int = 5
    n_aux: int = 0
    dtype: str = "f64"
    dissipation: str = "var0"
    output: str = "haloed"
    kernel: str = "auto"     # 'auto' | 'cell': 3-D shapes have a plane-marching kernel (auto) and a thread-per-cell one
    # 'reference': the reference's arithmetic, bit for bit (no contraction, IEEE division / sqrt).  'fast': permission to
    # contract multiply-adds and use a branch-free reciprocal / sqrt -- within 1e-12 relative, not bitwise; committed for
    # the headline shapes, other shapes keep the reference arithmetic (EXAHYPE_FLAG_FAST_ARITHMETIC)
    arithmetic: str = "reference"

    def __post_init__(self):
        from .KernelBuilder import viable
        if not viable(self.dim, self.patch_size, self.halo_size):
            raise Exception('check viability of inputs')          # reference KernelBuilder.py:52-53
        if self.model not in MODEL or self.dtype not in DTYPE:
            raise ValueError(f"unknown model/dtype {self.model}/{self.dtype}")
        if self.dissipation not in ("var0", "all") or self.output not in ("haloed", "unhaloed", "unknowns") \
                or self.kernel not in ("auto", "cell") or self.arithmetic not in ("reference", "fast"):
            raise ValueError("dissipation must be 'var0'|'all', output 'haloed'|'unhaloed'|'unknowns', kernel 'auto'|'cell', "
                             "arithmetic 'reference'|'fast'")
        self._lib = load()

    @classmethod
    def from_kernel(cls, kernel, model: str = "euler", **kw) -> "PatchUpdate":
        """Configuration taken from a :class:`KernelBuilder` declaration."""
        return cls(model=model, dim=kernel.dim, patch_size=kernel.patch_size, halo_size=kernel.halo_size,
                   n_real=kernel.n_real, n_aux=kernel.n_aux, **kw)

    # ------------------------------------------------------------------ geometry
    @property
    def n_var(self) -> int:
        return self.n_real + self.n_aux

    @property
    def side(self) -> int:
        return self.patch_size + 2 * self.halo_size

    @property
    def cells_per_patch(self) -> int:
        return self.patch_size ** self.dim

    def in_shape(self, n_patches: int):
        return (n_patches,) + (self.side,) * self.dim + (self.n_var,)

    def out_shape(self, n_patches: int):
        if self.output == "haloed":
            return self.in_shape(n_patches)
        return (n_patches,) + (self.patch_size,) * self.dim + (self.n_real if self.output == "unknowns" else self.n_var,)

    @property
    def algorithmic_bytes_per_patch(self) -> int:
        """read interior + face-halo cells once (all variables), write interior unknowns, one lambda."""
        P, d, h = self.patch_size, self.dim, self.halo_size
        need = P ** d + 2 * d * h * P ** (d - 1)
        es = 8 if self.dtype == "f64" else 4
        return es * (need * self.n_var + P ** d * self.n_real) + es

    def flags(self, accumulate_lambda: bool = False) -> int:
        return ((FLAG_DISSIPATION_ALL if self.dissipation == "all" else 0) |
                (FLAG_OUTPUT_UNHALOED if self.output in ("unhaloed", "unknowns") else 0) |
                (FLAG_OUTPUT_UNKNOWNS_ONLY if self.output == "unknowns" else 0) |
                (FLAG_LAMBDA_ACCUMULATE if accumulate_lambda else 0) |
                (FLAG_KERNEL_CELL if self.kernel == "cell" else 0) |
                (FLAG_FAST_ARITHMETIC if self.arithmetic == "fast" else 0))

    def config(self, accumulate_lambda: bool = False) -> FvConfig:
        return FvConfig(MODEL[self.model], DTYPE[self.dtype], self.dim, self.patch_size, self.halo_size,
                        self.n_real, self.n_aux, self.flags(accumulate_lambda))

    def supported(self) -> bool:
        c = self.config()
        return bool(self._lib.exahype_cuda_fv_supported(ctypes.byref(c)))

    def launch_info(self, n_patches: int) -> dict:
        c = self.config()
        g, b, s, t = (ctypes.c_int() for _ in range(4))
        check(self._lib.exahype_cuda_fv_launch_info(ctypes.byref(c), n_patches, g, b, s, t), self._lib)
        return dict(grid=g.value, block=b.value, smem_bytes=s.value, patches_per_tile=t.value)

    # ------------------------------------------------------------------ device-resident batch
    def _n_patches(self, numel: int) -> int:
        per = self.side ** self.dim * self.n_var
        if numel % per:
            raise ValueError(f"input holds {numel} values, not a whole number of {per}-value patches")
        return numel // per

    def step(self, q_in, q_out=None, dt: float = 0.0, lambda_patch=None, lambda_max=None, stream=None,
             accumulate_lambda: bool = False, reducer=None):
        """Asynchronous update of a batch on the current CUDA device.

        ``q_in``/``q_out``/``lambda_*`` are contiguous CUDA tensors of this object's dtype.  ``q_out=None`` means in
        place (haloed output only).  Returns ``q_out``.

        ``reducer`` (multi-GPU): a :class:`exahype_b200.dist.TimestepReducer` with the peer-memory backend -- the
        all-reduce(max) of ``lambda_max`` over the ranks then runs in the same launch where the kernel supports it
        (``exahype_cuda_fv_step_allreduce``); with any other reducer it is ``step`` followed by ``allreduce_max``.
        """
        import torch
        tdt = torch.float64 if self.dtype == "f64" else torch.float32
        if not q_in.is_cuda:
            raise ValueError("step() takes CUDA tensors; use time_step() for host arrays")
        if q_out is None:
            if self.output != "haloed":
                raise ValueError("un-haloed output needs an explicit q_out")
            q_out = q_in
        for name, t in (("q_in", q_in), ("q_out", q_out), ("lambda_patch", lambda_patch), ("lambda_max", lambda_max)):
            if t is None:
                continue
            if t.dtype != tdt or not t.is_contiguous() or not t.is_cuda or t.device != q_in.device:
                raise ValueError(f"{name} must be a contiguous {tdt} CUDA tensor on {q_in.device}")
        n = self._n_patches(q_in.numel())
        if q_out is not q_in and q_out.numel() != int(np.prod(self.out_shape(n))):
            raise ValueError(f"q_out must hold {self.out_shape(n)}")
        if lambda_patch is not None and lambda_patch.numel() < n:
            raise ValueError("lambda_patch must hold one value per patch")
        if stream is None:
            stream = torch.cuda.current_stream(q_in.device).cuda_stream
        c = self.config(accumulate_lambda)
        fused = reducer is not None and getattr(reducer, "peer_handle", None)
        if reducer is not None and lambda_max is None:
            raise ValueError("a reducer needs lambda_max")
        with torch.cuda.device(q_in.device):
            if fused:
                check(self._lib.exahype_cuda_fv_step_allreduce(
                    ctypes.byref(c), reducer.peer_handle, q_in.data_ptr(), q_out.data_ptr(), n, float(dt),
                    lambda_patch.data_ptr() if lambda_patch is not None else None, lambda_max.data_ptr(), stream),
                    self._lib)
            else:
                check(self._lib.exahype_cuda_fv_step(
                    ctypes.byref(c), q_in.data_ptr(), q_out.data_ptr(), n, float(dt),
                    lambda_patch.data_ptr() if lambda_patch is not None else None,
                    lambda_max.data_ptr() if lambda_max is not None else None, stream), self._lib)
        if reducer is not None and not fused:
            reducer.allreduce_max(lambda_max, stream=stream)
        return q_out

    def step_loop(self, loop, q_in, q_out, lambda_patch=None, stream=None):
        """One step of a device-resident time loop (:class:`exahype_b200.dist.TimeLoop`,
        ``exahype_cuda_fv_step_time_loop``): like :meth:`step`, but dt is the loop's -- ``cfl_dx / lambda_max`` of the
        previous step over all ranks, derived on the device -- and this step's ``lambda_max`` goes into the loop's
        exchange.  Collective across the loop's ranks.  ``q_in`` may be an empty shard."""
        import torch
        tdt = torch.float64 if self.dtype == "f64" else torch.float32
        if loop.dtype != self.dtype:
            raise ValueError(f"the loop runs in {loop.dtype}, this update in {self.dtype}")
        for name, t in (("q_in", q_in), ("q_out", q_out), ("lambda_patch", lambda_patch)):
            if t is None:
                continue
            if t.dtype != tdt or not t.is_contiguous() or not t.is_cuda or t.device != q_in.device:
                raise ValueError(f"{name} must be a contiguous {tdt} CUDA tensor on {q_in.device}")
        n = self._n_patches(q_in.numel())
        if q_out is not q_in and q_out.numel() != int(np.prod(self.out_shape(n))):
            raise ValueError(f"q_out must hold {self.out_shape(n)}")
        if q_out is q_in and self.output != "haloed":
            raise ValueError("un-haloed output needs its own q_out")
        if lambda_patch is not None and lambda_patch.numel() < n:
            raise ValueError("lambda_patch must hold one value per patch")
        if stream is None:
            stream = torch.cuda.current_stream(q_in.device).cuda_stream
        c = self.config()
        with torch.cuda.device(q_in.device):
            check(self._lib.exahype_cuda_fv_step_time_loop(
                ctypes.byref(c), loop.handle, q_in.data_ptr() if n else None, q_out.data_ptr() if n else None, n,
                lambda_patch.data_ptr() if lambda_patch is not None else None, stream), self._lib)
        return q_out

    def step_cell_data(self, q_in_ptrs, q_out_ptrs, dt=0.0, dt_patch=None, max_eigenvalue=None, lambda_max=None,
                       stream=None):
        """The ``CellData`` form (``exahype_cuda_fv_step_cell_data``): ``q_in_ptrs`` / ``q_out_ptrs`` are CUDA int64
        tensors of per-patch device pointers (``QIn`` haloed; ``QOut`` haloed or un-haloed per ``output``),
        ``dt_patch`` an optional CUDA tensor of per-patch time steps, ``max_eigenvalue`` an optional per-patch output."""
        import torch
        tdt = torch.float64 if self.dtype == "f64" else torch.float32
        n = int(q_in_ptrs.numel())
        for name, t, want in (("q_in_ptrs", q_in_ptrs, torch.int64), ("q_out_ptrs", q_out_ptrs, torch.int64),
                              ("dt_patch", dt_patch, tdt), ("max_eigenvalue", max_eigenvalue, tdt),
                              ("lambda_max", lambda_max, tdt)):
            if t is None:
                continue
            if not t.is_cuda or not t.is_contiguous() or t.dtype != want:
                raise ValueError(f"{name} must be a contiguous CUDA tensor of dtype {want}")
            if name not in ("lambda_max",) and t.numel() < n:
                raise ValueError(f"{name} must hold one value per patch")
        if stream is None:
            stream = torch.cuda.current_stream(q_in_ptrs.device).cuda_stream
        cells = CellData(n, q_in_ptrs.data_ptr(), q_out_ptrs.data_ptr(),
                         dt_patch.data_ptr() if dt_patch is not None else None,
                         max_eigenvalue.data_ptr() if max_eigenvalue is not None else None, None, None, None)
        c = self.config()
        with torch.cuda.device(q_in_ptrs.device):
            check(self._lib.exahype_cuda_fv_step_cell_data(
                ctypes.byref(c), ctypes.byref(cells), float(dt),
                lambda_max.data_ptr() if lambda_max is not None else None, stream), self._lib)

    def fill_synthetic(self, q, first_patch: int = 0, seed: int = 20240601, stream=None):
        """Fills the CUDA tensor ``q`` (``in_shape(n)``) with the benchmark's synthetic admissible state for the global
        patches ``first_patch .. first_patch+n``, on the device."""
        import torch
        n = self._n_patches(q.numel())
        if not q.is_cuda or not q.is_contiguous() or q.dtype != (torch.float64 if self.dtype == "f64" else torch.float32):
            raise ValueError("q must be a contiguous CUDA tensor of this object's dtype")
        if stream is None:
            stream = torch.cuda.current_stream(q.device).cuda_stream
        cells = self.side ** self.dim
        c = self.config()
        with torch.cuda.device(q.device):
            check(self._lib.exahype_cuda_fill_synthetic(ctypes.byref(c), q.data_ptr(), first_patch * cells, n * cells,
                                                        seed, stream), self._lib)
        return q

    # ------------------------------------------------------------------ the reference's call shape, host memory
    def time_step(self, Q: np.ndarray, dt: float, Q_out: Optional[np.ndarray] = None, lambda_patch=None):
        """``time_step(Q, dt)`` on host memory (numpy array or pinned CPU tensor's ``.numpy()``): ``Q`` is updated in
        place, or ``Q_out`` receives the result for un-haloed output.  Returns the batch's max eigenvalue."""
        npdt = _np_dtype(self.dtype)
        if Q.dtype != npdt or not Q.flags["C_CONTIGUOUS"]:
            raise ValueError(f"Q must be a C-contiguous {npdt.__name__} array")
        n = self._n_patches(Q.size)
        if Q_out is None:
            if self.output != "haloed":
                raise ValueError("un-haloed output needs Q_out")
            Q_out = Q
        elif Q_out.dtype != npdt or not Q_out.flags["C_CONTIGUOUS"] or Q_out.size != int(np.prod(self.out_shape(n))):
            raise ValueError(f"Q_out must be a C-contiguous {npdt.__name__} array of shape {self.out_shape(n)}")
        lam_max = np.zeros(1, dtype=npdt)
        c = self.config()
        check(self._lib.exahype_cuda_time_step_host(
            ctypes.byref(c), Q.ctypes.data, Q_out.ctypes.data, n, float(dt),
            lambda_patch.ctypes.data if lambda_patch is not None else None, lam_max.ctypes.data), self._lib)
        return lam_max[0]


def time_step(Q: np.ndarray, dt: float, *, dim: int, patch_size: int, halo_size: int = 1, n_real: int, n_aux: int = 0,
              model: str = "euler", dissipation: str = "var0"):
    """Functional form of the reference's ``void time_step(double* Q, double dt)`` for a host batch."""
    dtype = "f64" if Q.dtype == np.float64 else "f32"
    return PatchUpdate(model=model, dim=dim, patch_size=patch_size, halo_size=halo_size, n_real=n_real, n_aux=n_aux,
                       dtype=dtype, dissipation=dissipation).time_step(Q, dt)
