"""One process per GPU: how the hot path shards and the collectives it needs.

* feature build: scaffolds are split into contiguous ranges of about equal base pairs (`shard_scaffolds`); every rank
  builds the rows of its scaffolds with no exchange; the row blocks are all-gathered so that every rank holds all
  datapoints (`allgather_rows`).
* split search: dimensions are split into contiguous blocks (`dim_block`); `abw_search_run_sharded` needs an
  all-gather of per-cluster best records and a sum in which only the owner of the winning dimension contributes
  non-zeros.  `TorchCollectives` implements both with torch.distributed (NCCL over NVLink on GPUs).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi


def shard_scaffolds(lengths, world):
    """Contiguous scaffold ranges [lo, hi) per rank with about equal total length (row order is preserved)."""
    lengths = np.asarray(lengths, dtype=np.int64)
    cum = np.concatenate([[0], np.cumsum(lengths)])
    total = int(cum[-1])
    bounds = [0]
    for r in range(1, world):
        bounds.append(int(np.searchsorted(cum, total * r / world, side="left")))
    bounds.append(lengths.size)
    bounds = np.maximum.accumulate(np.clip(bounds, 0, lengths.size))
    return [(int(bounds[r]), int(bounds[r + 1])) for r in range(world)]


def dim_block(D, rank, world):
    """Contiguous block of dimensions owned by `rank`: (offset, count); blocks differ by at most one dimension."""
    base, extra = divmod(D, world)
    off = rank * base + min(rank, extra)
    return off, base + (1 if rank < extra else 0)


class _DevArray:
    """Minimal __cuda_array_interface__ wrapper so that torch can view a raw device pointer without copying."""

    def __init__(self, ptr, nbytes, typestr="|u1", itemsize=1):
        self.__cuda_array_interface__ = {"shape": (nbytes // itemsize,), "typestr": typestr, "data": (int(ptr), False), "version": 2}


class TorchCollectives:
    """abw_collectives backed by torch.distributed (backend nccl on GPUs).  The callbacks are host-synchronous."""

    def __init__(self, device_index):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.device = torch.device("cuda", device_index)
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.calls = 0
        self.bytes = 0

        views = {}                                         # (pointer, bytes, dtype) -> tensor view: the search reuses the same device buffers level after level

        def view(ptr, nbytes, typestr="|u1", itemsize=1):
            key = (ptr, nbytes, typestr)
            t = views.get(key)
            if t is None:
                if len(views) > 256:
                    views.clear()
                t = views[key] = torch.as_tensor(_DevArray(ptr, nbytes, typestr, itemsize), device=self.device)
            return t

        def allgather(user, d_send, d_recv, nbytes):
            try:
                send = view(d_send, nbytes)
                recv = view(d_recv, nbytes * self.world)
                dist.all_gather_into_tensor(recv, send)
                torch.cuda.synchronize(self.device)
                self.calls += 1
                self.bytes += nbytes * self.world
                return 0
            except Exception as e:  # never let an exception cross the C boundary
                print("allgather callback failed:", repr(e), flush=True)
                return 1

        def allreduce(user, d_buf, count):
            try:
                t = view(d_buf, count * 8, "<i8", 8)
                dist.all_reduce(t, op=dist.ReduceOp.SUM)
                torch.cuda.synchronize(self.device)
                self.calls += 1
                self.bytes += count * 8
                return 0
            except Exception as e:
                print("allreduce callback failed:", repr(e), flush=True)
                return 1

        self._ag, self._ar = capi.ALLGATHER_FN(allgather), capi.ALLREDUCE_FN(allreduce)
        self.struct = capi.Collectives(self._ag, self._ar, None, self.rank, self.world)


class NcclCollectives:
    """abw_collectives implemented inside libabawaca_b200.so on NCCL (abw_nccl_collectives_create): the operations are enqueued on the context
    stream, nothing returns to the interpreter during a search.  torch.distributed only carries the 128-byte NCCL id to the other ranks."""

    def __init__(self, ctx, device_index):
        import torch
        import torch.distributed as dist
        self.ctx = ctx
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        idbuf = (C.c_ubyte * 128)()
        if self.rank == 0:
            ctx.check(ctx.lib.abw_nccl_unique_id(ctx.h, idbuf))
        t = torch.tensor(list(bytes(idbuf)), dtype=torch.uint8, device=torch.device("cuda", device_index))
        dist.broadcast(t, src=0)
        idbytes = bytes(t.cpu().tolist())
        self.struct = capi.Collectives()
        ctx.check(ctx.lib.abw_nccl_collectives_create(ctx.h, idbytes, self.rank, self.world, C.byref(self.struct)))
        self.calls = self.bytes = None                     # not counted: no callback into the interpreter

    def close(self):
        if self.struct.user:
            self.ctx.lib.abw_nccl_collectives_destroy(C.byref(self.struct))


def allgather_rows(torch, dist, local_rows, counts):
    """local_rows: device tensor [n_local][ncols] (float64); counts: rows per rank.  Returns [sum(counts)][ncols] on every rank."""
    world = len(counts)
    nmax = max(counts)
    ncols = local_rows.shape[1]
    pad = torch.zeros((nmax, ncols), dtype=local_rows.dtype, device=local_rows.device)
    pad[: local_rows.shape[0]] = local_rows
    out = torch.empty((world * nmax, ncols), dtype=local_rows.dtype, device=local_rows.device)
    dist.all_gather_into_tensor(out, pad)
    if all(c == nmax for c in counts):
        return out
    return torch.cat([out[r * nmax: r * nmax + counts[r]] for r in range(world)], dim=0)
