"""Multi-GPU host logic: patch sharding and the global admissible-time-step reduction.

The reference has no distributed code.  Patches are independent given their halos, so the
batch axis shards with no data-path collective; the only exchange is one ``allreduce(max)`` of a single scalar -- the
largest eigenvalue -- per step, from which every rank derives the same ``dt = CFL * dx / lambda_max``.  ``max`` is exact,
so an N-GPU run equals the 1-GPU run bit for bit.

One process per GPU.  ``torch.distributed`` is the plumbing (rendezvous, broadcasting the NCCL unique id); the
reduction itself is ``ncclAllReduce`` issued by ``libexahype_cuda.so`` on the caller's stream
(``exahype_cuda_allreduce_max``), so it is stream-ordered behind the patch-update kernel without a host sync.  On CPU
tensors (tests, ``gloo``) the same class reduces through ``torch.distributed``.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass


@dataclass(frozen=True)
class PatchSharding:
    """Rank ``r`` of ``R`` owns the contiguous patch range ``[floor(r*B/R), floor((r+1)*B/R))``."""
    global_patches: int
    world_size: int
    rank: int

    def __post_init__(self):
        if self.world_size < 1 or not (0 <= self.rank < self.world_size) or self.global_patches < 0:
            raise ValueError("bad sharding")

    @property
    def first(self) -> int:
        return self.rank * self.global_patches // self.world_size

    @property
    def last(self) -> int:
        return (self.rank + 1) * self.global_patches // self.world_size

    @property
    def count(self) -> int:
        return self.last - self.first

    def slice(self):
        return slice(self.first, self.last)


def admissible_dt(lambda_max: float, cell_size: float, cfl: float = 0.9) -> float:
    """``dt = CFL * dx / lambda_max`` for the next step (0 eigenvalue -> no constraint)."""
    return float("inf") if lambda_max <= 0.0 else cfl * cell_size / lambda_max


class TimestepReducer:
    """All-reduce(max) of the per-GPU largest eigenvalue.

    CUDA tensors, ``backend``:
      ``'peer'``  one-shot exchange over NVLink peer memory (``exahype_cuda_peer_reducer_*``: every rank's mailbox is
                  mapped into every peer through CUDA IPC; one tiny stream-ordered kernel per step) -- the low-latency path;
      ``'nccl'``  ``ncclAllReduce`` through ``libexahype_cuda.so`` (own communicator);
      ``'auto'``  ``'peer'`` when every rank could map every mailbox (one node, P2P capable), else ``'nccl'``.
    The handles / the NCCL unique id travel over the ``torch.distributed`` default group.  Both give the same bits (max is
    exact).  CPU tensors (tests, ``gloo``): ``torch.distributed.all_reduce(MAX)`` on the default group.
    """

    def __init__(self, world_size: int, rank: int, use_nccl: bool = True, backend: str = "auto"):
        if backend not in ("auto", "peer", "nccl"):
            raise ValueError("backend must be 'auto', 'peer' or 'nccl'")
        self.world_size, self.rank = world_size, rank
        self._comm = ctypes.c_void_p()
        self._peer = ctypes.c_void_p()
        self._lib = None
        self.backend = "none"
        if not use_nccl or world_size <= 1:
            return
        import torch
        from . import runtime
        self._lib = runtime.load()
        with torch.cuda.device(torch.cuda.current_device()):
            if backend in ("auto", "peer") and self._connect_peers():
                self.backend = "peer"
                return
            if backend == "peer":
                raise RuntimeError("peer-memory reducer unavailable: " + self._peer_error)
            self._init_nccl()
            self.backend = "nccl"

    def _all_ok(self, ok: bool) -> bool:
        import torch.distributed as dist
        flags = [None] * self.world_size
        dist.all_gather_object(flags, bool(ok))
        return all(flags)

    def _connect_peers(self) -> bool:
        import torch.distributed as dist
        lib = self._lib
        self._peer_error = ""
        handle = (ctypes.c_char * 64)()
        ok = lib.exahype_cuda_peer_reducer_create(ctypes.byref(self._peer), self.world_size, self.rank) == 0 and \
            lib.exahype_cuda_peer_reducer_local_handle(self._peer, handle) == 0
        if not ok:
            self._peer_error = lib.exahype_cuda_last_error().decode()
        gathered = [None] * self.world_size
        dist.all_gather_object(gathered, bytes(handle) if ok else None)
        if all(g is not None for g in gathered):
            blob = (ctypes.c_char * (64 * self.world_size)).from_buffer_copy(b"".join(gathered))
            ok = lib.exahype_cuda_peer_reducer_connect(self._peer, blob) == 0
            if not ok:
                self._peer_error = lib.exahype_cuda_last_error().decode()
        else:
            ok = False
        if self._all_ok(ok):
            return True
        if self._peer:
            lib.exahype_cuda_peer_reducer_destroy(self._peer)
            self._peer = ctypes.c_void_p()
        return False

    def _init_nccl(self):
        import torch.distributed as dist
        from . import runtime
        ident = (ctypes.c_char * 128)()
        if self.rank == 0:
            runtime.check(self._lib.exahype_cuda_nccl_unique_id(ident), self._lib)
        payload = [bytes(ident)]
        dist.broadcast_object_list(payload, src=0)
        ident = (ctypes.c_char * 128).from_buffer_copy(payload[0])
        runtime.check(self._lib.exahype_cuda_comm_init(ctypes.byref(self._comm), ident, self.world_size, self.rank),
                      self._lib)

    def allreduce_max(self, value, stream=None):
        """In place on ``value`` (1-element tensor; small tensors with the NCCL backend); asynchronous on ``stream`` for
        CUDA tensors."""
        if self.world_size == 1:
            return value
        import torch
        import torch.distributed as dist
        if value.is_cuda and (self._peer or self._comm):
            from . import runtime
            if stream is None:
                stream = torch.cuda.current_stream(value.device).cuda_stream
            dtype = {torch.float64: 0, torch.float32: 1}[value.dtype]
            if self._peer:
                if value.numel() != 1:
                    raise ValueError("the peer-memory reducer exchanges one scalar per step")
                runtime.check(self._lib.exahype_cuda_peer_reducer_allreduce_max(self._peer, value.data_ptr(), dtype, stream),
                              self._lib)
            else:
                runtime.check(self._lib.exahype_cuda_allreduce_max(self._comm, value.data_ptr(), value.numel(), dtype,
                                                                   stream), self._lib)
        else:
            dist.all_reduce(value, op=dist.ReduceOp.MAX)
        return value

    @property
    def peer_handle(self):
        """The library's peer-memory reducer (None unless ``backend == 'peer'``): what ``PatchUpdate.step(reducer=...)``
        hands to ``exahype_cuda_fv_step_allreduce`` for the all-reduce in the patch kernel's own epilogue."""
        return self._peer if self._peer else None

    def timed_out(self) -> bool:
        """True if a peer-memory wait gave up (a rank never arrived): the waiting side's result is NaN and every later
        call on this reducer raises ``TIMEOUT``.  The flag lives in host-mapped memory -- nothing is synchronised."""
        if not self._peer:
            return False
        flag = ctypes.c_int(0)
        self._lib.exahype_cuda_peer_reducer_status(self._peer, ctypes.byref(flag))
        return bool(flag.value)

    def set_timeout(self, seconds: float):
        """How long a peer-memory wait spins before it gives up (default 10 s)."""
        if self._peer:
            from . import runtime
            runtime.check(self._lib.exahype_cuda_peer_reducer_set_timeout(self._peer, float(seconds)), self._lib)

    def enable_trace(self, capacity: int = 1024):
        """Keep device-side ``globaltimer`` stamps of the last ``capacity`` exchanges (csrc/peer_mail.cuh FV_TRACE_*)."""
        if self._peer:
            from . import runtime
            runtime.check(self._lib.exahype_cuda_peer_reducer_enable_trace(self._peer, int(capacity)), self._lib)

    def read_trace(self, first_seq: int, count: int):
        """``[count, 8]`` uint64 array of the stamps of exchanges ``first_seq ..`` (synchronises the device)."""
        import numpy as np
        from . import runtime
        out = np.zeros((count, 8), dtype=np.uint64)
        runtime.check(self._lib.exahype_cuda_peer_reducer_read_trace(self._peer, first_seq, count, out.ctypes.data), self._lib)
        return out

    def close(self):
        if self._lib is None:
            return
        if self._peer:
            self._lib.exahype_cuda_peer_reducer_destroy(self._peer)
            self._peer = ctypes.c_void_p()
        if self._comm:
            self._lib.exahype_cuda_comm_destroy(self._comm)
            self._comm = ctypes.c_void_p()


class TimeLoop:
    """Device-resident time loop (``exahype_cuda_time_loop_*``): step ``k+1`` advances with
    ``dt = cfl_dx / max over all ranks of lambda_max(step k)``, produced and consumed on the device -- no host round
    trip, no memset and (for the warp-per-patch kernel) no second kernel between two steps.  The exchange is
    split-phase: step ``k``'s last warp publishes into every peer's mailbox, step ``k+1``'s warps consume.

    ``reducer``: a :class:`TimestepReducer` with the peer-memory backend (all its ranks step in lockstep) or ``None``
    for one GPU.  Drive it with :meth:`exahype_b200.runtime.PatchUpdate.step_loop`.
    """

    def __init__(self, dtype: str = "f64", reducer: "TimestepReducer | None" = None, cfl_dx: float = 1.0,
                 dt0: float = 0.0, history_capacity: int = 0, peer_handle=None):
        from . import runtime
        self._lib = runtime.load()
        self.dtype = dtype
        self.cfl_dx, self.dt0 = float(cfl_dx), float(dt0)
        self.handle = ctypes.c_void_p()
        if reducer is not None and peer_handle is None:
            peer_handle = reducer.peer_handle
            if peer_handle is None and reducer.world_size > 1:
                raise RuntimeError("the device-resident time loop needs the peer-memory reducer (backend='peer')")
        self._reducer = reducer
        runtime.check(self._lib.exahype_cuda_time_loop_create(ctypes.byref(self.handle), runtime.DTYPE[dtype], peer_handle,
                                                              self.cfl_dx, self.dt0, int(history_capacity)), self._lib)

    @property
    def steps(self) -> int:
        return int(self._lib.exahype_cuda_time_loop_steps(self.handle))

    def flush(self, stream=None):
        """Consume the exchange still in flight: afterwards :meth:`dt_device` holds the NEXT step's dt."""
        import torch
        from . import runtime
        if stream is None:
            stream = torch.cuda.current_stream().cuda_stream
        runtime.check(self._lib.exahype_cuda_time_loop_flush(self.handle, stream), self._lib)

    def history(self, first: int = 0, count: "int | None" = None):
        """``[count, 4]`` array: per step ``{dt used, global lambda_max it was derived from (0: dt0), this device's
        lambda_max of the step, 0}``; entry ``steps`` (after :meth:`flush`) is ``{next dt, last global lambda_max}``.
        Synchronises the device."""
        import numpy as np
        from . import runtime
        if count is None:
            count = self.steps - first
        out = np.zeros((count, 4), dtype=np.float64 if self.dtype == "f64" else np.float32)
        runtime.check(self._lib.exahype_cuda_time_loop_history(self.handle, first, count, out.ctypes.data), self._lib)
        return out

    def dt_device(self) -> int:
        """Device address of the scalar holding the most recently established time step."""
        from . import runtime
        p = ctypes.c_void_p()
        runtime.check(self._lib.exahype_cuda_time_loop_dt_device(self.handle, ctypes.byref(p)), self._lib)
        return p.value

    def close(self):
        if self.handle:
            self._lib.exahype_cuda_time_loop_destroy(self.handle)
            self.handle = ctypes.c_void_p()


class LocalPeerGroup:
    """``world_size`` peer-memory reducers living in THIS process on the current device, connected to each other
    through plain device pointers (``exahype_cuda_peer_reducer_connect_local``).  The same mailbox protocol and the
    same kernels as one process per GPU, minus NVLink: what the 1-GPU tests drive (each "rank" on its own stream), and
    the way to run several ranks per GPU.  ``group[r]`` quacks like a :class:`TimestepReducer` with the peer backend."""

    class Rank:
        backend = "peer"

        def __init__(self, group, rank, handle):
            self._group, self.rank, self.world_size, self._peer, self._lib = group, rank, group.world_size, handle, group._lib

        @property
        def peer_handle(self):
            return self._peer

        def allreduce_max(self, value, stream=None):
            import torch
            from . import runtime
            if stream is None:
                stream = torch.cuda.current_stream(value.device).cuda_stream
            dtype = {torch.float64: 0, torch.float32: 1}[value.dtype]
            runtime.check(self._lib.exahype_cuda_peer_reducer_allreduce_max(self._peer, value.data_ptr(), dtype, stream),
                          self._lib)
            return value

        timed_out = TimestepReducer.timed_out
        set_timeout = TimestepReducer.set_timeout
        enable_trace = TimestepReducer.enable_trace
        read_trace = TimestepReducer.read_trace

    def __init__(self, world_size: int):
        from . import runtime
        self._lib = runtime.load()
        self.world_size = world_size
        handles = (ctypes.c_void_p * world_size)()
        self.ranks = []
        for r in range(world_size):
            h = ctypes.c_void_p()
            runtime.check(self._lib.exahype_cuda_peer_reducer_create(ctypes.byref(h), world_size, r), self._lib)
            handles[r] = h.value
            self.ranks.append(LocalPeerGroup.Rank(self, r, h))
        runtime.check(self._lib.exahype_cuda_peer_reducer_connect_local(handles, world_size), self._lib)

    def __getitem__(self, r):
        return self.ranks[r]

    def __len__(self):
        return self.world_size

    def close(self):
        for rk in self.ranks:
            if rk._peer:
                self._lib.exahype_cuda_peer_reducer_destroy(rk._peer)
                rk._peer = ctypes.c_void_p()
