"""Cell sharding across GPUs (one process per GPU, torch.distributed plumbing).

Cells are conditionally independent given the parameters, so every rank owns a
contiguous range of cells (columns of the ``G x n`` matrices, rows of ``z``)
and the only exchange is a sum of the M-step sufficient statistics: one
all-reduce per EM iteration plus two at call start.  The
reference has no multi-process path at all (only ``foreach %dopar%`` over
genes, R/EM_impute.R:238); this replaces it.

The CUDA library calls back into ``CellShard.allreduce_sum_`` with a raw
device pointer; on GPU it is wrapped zero-copy into a torch tensor and handed
to NCCL.  With the gloo backend (CPU tests) the same code path runs on host
pointers.
"""
from __future__ import annotations

import ctypes as C

import numpy as np


def partition(n_total: int, world_size: int):
    """Contiguous, balanced cell ranges: boundaries b[0..world_size]."""
    base, rem = divmod(int(n_total), int(world_size))
    b = [0]
    for r in range(world_size):
        b.append(b[-1] + base + (1 if r < rem else 0))
    return b


def _is_device_pointer(ptr):
    """True when `ptr` is CUDA device memory (cudaPointerGetAttributes through cuda-python / torch); host otherwise."""
    try:
        import torch
        if not torch.cuda.is_available():
            return False
        from cuda import cudart            # cuda-python
        err, at = cudart.cudaPointerGetAttributes(int(ptr))
        return int(err) == 0 and int(at.type) in (2, 3)     # cudaMemoryTypeDevice / Managed
    except Exception:  # noqa: BLE001
        return False


class _DevPtr:
    """Expose a raw CUDA pointer through __cuda_array_interface__."""

    def __init__(self, ptr, count):
        self.__cuda_array_interface__ = {"shape": (int(count),), "typestr": "<f8", "data": (int(ptr), False),
                                         "version": 3, "strides": None}


class CellShard:
    """This rank's shard of the cells + the all-reduce used by the library."""

    def __init__(self, n_total, rank=None, world_size=None, group=None, native=False):
        """``native=True``: the library reduces over its own NCCL rank communicator (built here from a unique id
        broadcast over torch.distributed); no Python callback runs inside the EM loop."""
        import torch.distributed as dist
        self._dist = dist
        self.group = group
        self.native = bool(native)
        if self.native:
            from ._lib import comm_init_from_torch, load
            if load().simples_comm_nranks() != dist.get_world_size(group):
                comm_init_from_torch(group)
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world_size = dist.get_world_size(group) if world_size is None else world_size
        self.bounds = partition(n_total, self.world_size)
        self.n_total = int(n_total)
        self.cell_offset = self.bounds[self.rank]
        self.n_local = self.bounds[self.rank + 1] - self.bounds[self.rank]
        self.n_allreduce = 0
        self.bytes_allreduce = 0

    # ---- slicing helpers (R orientation: cells are columns of G x n, rows of n x M0) ----
    def cols(self, A):
        return np.asfortranarray(np.asarray(A)[:, self.cell_offset:self.cell_offset + self.n_local])

    def rows(self, A):
        return np.asarray(A)[self.cell_offset:self.cell_offset + self.n_local]

    # ---- collective ----
    def allreduce_sum_(self, ptr, count, stream=None):
        import torch
        dist = self._dist
        self.n_allreduce += 1
        self.bytes_allreduce += int(count) * 8
        if dist.get_backend(self.group) == "nccl":
            t = torch.as_tensor(_DevPtr(ptr, count), device="cuda")
            ext = torch.cuda.ExternalStream(int(stream)) if stream else torch.cuda.current_stream()
            with torch.cuda.stream(ext):
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        elif _is_device_pointer(ptr):
            # gloo group but the CUDA library is calling (it always passes DEVICE pointers): stage through the host,
            # ordered on the library's stream
            t = torch.as_tensor(_DevPtr(ptr, count), device="cuda")
            ext = torch.cuda.ExternalStream(int(stream)) if stream else torch.cuda.current_stream()
            with torch.cuda.stream(ext):
                h = t.cpu()
                dist.all_reduce(h, op=dist.ReduceOp.SUM, group=self.group)
                t.copy_(h)
                ext.synchronize()
        else:  # gloo: host pointer (CPU tests of the sharding logic)
            buf = (C.c_double * int(count)).from_address(int(ptr))
            t = torch.from_numpy(np.frombuffer(buf, dtype=np.float64))
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)

    def _host_reduce(self, x, op):
        import torch
        dist = self._dist
        t = torch.from_numpy(np.array(x, dtype=np.float64))
        if dist.get_backend(self.group) == "nccl":
            t = t.cuda()
        dist.all_reduce(t, op=op, group=self.group)
        return t.cpu().numpy()

    def sum_host(self, x):
        """Sum of a small host array over the ranks (driver bookkeeping: gene expression rates, cluster sizes)."""
        return self._host_reduce(x, self._dist.ReduceOp.SUM)

    def max_host(self, x):
        return self._host_reduce(x, self._dist.ReduceOp.MAX)

    def gather_cols(self, A_local):
        """Gather a G x n_local matrix from every rank into G x n_total (all ranks)."""
        import torch
        dist = self._dist
        parts = [None] * self.world_size
        dist.all_gather_object(parts, np.ascontiguousarray(A_local), group=self.group)
        return np.concatenate(parts, axis=1)


class Replicas:
    """One process per GPU, each holding the WHOLE data set (no cell sharding): independent runs are dealt to the
    ranks and only small Python objects are exchanged (``selectKM(..., replicas=Replicas())``)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self._dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world_size = dist.get_world_size(group)

    def allgather(self, obj):
        parts = [None] * self.world_size
        self._dist.all_gather_object(parts, obj, group=self.group)
        return parts
