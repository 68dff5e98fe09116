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


class PeerExchange:
    """The column exchange between the feature stage and the dimension-sharded search over NVLink peer memory (csrc/peer.cu): every rank owns one
    exchange buffer (two halves, used alternately), exported through CUDA IPC and mapped by all other ranks; scatter() converts the local rows to integer
    thousandths and stores every rank's columns (rank, rank + world, ...) straight into that rank's buffer, and barrier() -- a one-word stream-ordered
    all-reduce of the search's own collectives -- tells the receivers that everybody's rows have arrived.
    ok is False when CUDA IPC is not available on some rank (decided collectively): the caller keeps its NCCL all-to-all."""

    def __init__(self, ctx, torch, dist, device_index, half_bytes, coll):
        self.ctx, self.torch, self.dist, self.coll = ctx, torch, dist, coll
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.dev = torch.device("cuda", device_index)
        self.half_bytes = (int(half_bytes) + 255) & ~255
        self.buf = C.c_void_p()
        self.group = C.c_void_p()
        self.parity = 0
        handle = (C.c_ubyte * 64)()
        rc = ctx.lib.abw_peer_buffer_create(ctx.h, 2 * self.half_bytes, C.byref(self.buf), handle)
        self.why = "" if rc == 0 else "abw_peer_buffer_create: " + ctx.lib.abw_last_error(ctx.h).decode()
        mine = torch.tensor(list(bytes(handle)) + [1 if rc == 0 else 0], dtype=torch.uint8, device=self.dev)
        everybody = torch.empty(self.world * 65, dtype=torch.uint8, device=self.dev)
        dist.all_gather_into_tensor(everybody, mine)
        everybody = everybody.cpu().numpy().reshape(self.world, 65)
        ok = bool(everybody[:, 64].all())
        if ok:
            handles = np.ascontiguousarray(everybody[:, :64]).tobytes()
            ok = ctx.lib.abw_peer_group_create(ctx.h, self.buf, handles, self.rank, self.world, C.byref(self.group)) == 0
            if not ok:
                self.why = "abw_peer_group_create: " + ctx.lib.abw_last_error(ctx.h).decode()
        verdict = torch.tensor([1 if ok else 0], dtype=torch.int32, device=self.dev)
        dist.all_reduce(verdict, op=dist.ReduceOp.MIN)
        self.ok = bool(int(verdict.item()))
        self.word = ctx.alloc(64)
        ctx.memset(self.word, 0, 64)
        if not self.ok:
            import sys
            print(f"PeerExchange (rank {self.rank}): peer memory is not available, {self.why or 'another rank failed'}", file=sys.stderr, flush=True)
            self.close()

    def scatter(self, d_rows, nrows, ld, ncols, row0, d_inexact=None, segs=None):
        """enqueue the scatter of this rank's rows into the current half of everybody's buffer; returns the device pointer of MY matrix in that half.
        segs (the abw_segments the rows belong to): rows of scaffolds with a single window are left out; row0 counts kept rows."""
        off = self.parity * self.half_bytes
        self.ctx.check(self.ctx.lib.abw_scatter_columns_milli(self.ctx.h, self.group, segs, C.c_void_p(d_rows), nrows, ld, ncols, row0, off, C.c_void_p(d_inexact or 0)))
        self.parity ^= 1
        return self.buf.value + off

    def barrier(self):
        """stream ordered: what follows on the context stream sees the stores of every rank's scatter that was enqueued before its barrier"""
        st = self.coll.struct
        if st.allreduce_sum_i64(st.user, self.word, 1) != 0:
            raise RuntimeError("PeerExchange.barrier: the collective failed")

    def close(self):
        if self.group:
            self.ctx.lib.abw_peer_group_destroy(self.group)
            self.group = C.c_void_p()
        if self.buf:
            self.ctx.lib.abw_peer_buffer_destroy(self.ctx.h, self.buf)
            self.buf = C.c_void_p()


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
