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

    CUDA tensors: NCCL through ``libexahype_cuda.so`` (own communicator; the unique id travels over the
    ``torch.distributed`` default group).  CPU tensors: ``torch.distributed.all_reduce(MAX)`` on the default group.
    """

    def __init__(self, world_size: int, rank: int, use_nccl: bool = True):
        self.world_size, self.rank = world_size, rank
        self._comm = ctypes.c_void_p()
        self._lib = None
        if use_nccl and world_size > 1:
            import torch
            import torch.distributed as dist
            from . import runtime
            self._lib = runtime.load()
            ident = (ctypes.c_char * 128)()
            if rank == 0:
                runtime.check(self._lib.exahype_cuda_nccl_unique_id(ident), self._lib)
            payload = [bytes(ident)]
            dist.broadcast_object_list(payload, src=0)
            ident = (ctypes.c_char * 128).from_buffer_copy(payload[0])
            with torch.cuda.device(torch.cuda.current_device()):
                runtime.check(self._lib.exahype_cuda_comm_init(ctypes.byref(self._comm), ident, world_size, rank),
                              self._lib)

    def allreduce_max(self, value, stream=None):
        """In place on ``value`` (1-element or small tensor); asynchronous on ``stream`` for CUDA tensors."""
        if self.world_size == 1:
            return value
        import torch
        import torch.distributed as dist
        if value.is_cuda and self._comm:
            from . import runtime
            if stream is None:
                stream = torch.cuda.current_stream(value.device).cuda_stream
            dtype = {torch.float64: 0, torch.float32: 1}[value.dtype]
            runtime.check(self._lib.exahype_cuda_allreduce_max(self._comm, value.data_ptr(), value.numel(), dtype,
                                                               stream), self._lib)
        else:
            dist.all_reduce(value, op=dist.ReduceOp.MAX)
        return value

    def close(self):
        if self._comm and self._lib is not None:
            self._lib.exahype_cuda_comm_destroy(self._comm)
            self._comm = ctypes.c_void_p()
