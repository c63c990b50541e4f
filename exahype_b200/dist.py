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
        """True if a peer-memory wait gave up (a rank never arrived).  Synchronises the device."""
        if not self._peer:
            return False
        flag = ctypes.c_int(0)
        self._lib.exahype_cuda_peer_reducer_status(self._peer, ctypes.byref(flag))
        return bool(flag.value)

    def close(self):
        if self._lib is None:
            return
        if self._peer:
            self._lib.exahype_cuda_peer_reducer_destroy(self._peer)
            self._peer = ctypes.c_void_p()
        if self._comm:
            self._lib.exahype_cuda_comm_destroy(self._comm)
            self._comm = ctypes.c_void_p()
