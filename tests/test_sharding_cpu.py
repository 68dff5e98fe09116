"""CPU tests (gloo, world_size 2) of the host-side sharding logic used at N > 1: scaffold ranges, dimension blocks and the
padded all-gather of feature rows.  The CUDA kernels themselves are exercised by the -m gpu tests."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from abawaca_b200 import distributed


def test_shard_scaffolds_covers_everything_in_order():
    rng = np.random.default_rng(0)
    lengths = rng.integers(4000, 200000, 1000)
    for world in (1, 2, 3, 8):
        sh = distributed.shard_scaffolds(lengths, world)
        assert sh[0][0] == 0 and sh[-1][1] == lengths.size
        assert all(a[1] == b[0] for a, b in zip(sh, sh[1:]))
        loads = [lengths[lo:hi].sum() for lo, hi in sh]
        assert max(loads) - min(loads) <= 2 * lengths.max()


def test_dim_blocks_partition_the_dimensions():
    for D in (1, 7, 182, 189, 229):
        for world in (1, 2, 4, 8):
            blocks = [distributed.dim_block(D, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and sum(c for _, c in blocks) == D
            assert all(blocks[r][0] + blocks[r][1] == blocks[r + 1][0] for r in range(world - 1))
            assert max(c for _, c in blocks) - min(c for _, c in blocks) <= 1


def _worker(rank, world, port, counts, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ncols = 5
    start = sum(counts[:rank])
    local = torch.arange(start * ncols, (start + counts[rank]) * ncols, dtype=torch.float64).reshape(counts[rank], ncols)
    full = distributed.allgather_rows(torch, dist, local, counts)
    q.put((rank, full.numpy().copy()))
    dist.destroy_process_group()


@pytest.mark.parametrize("counts", [[4, 4], [3, 6]])
def test_allgather_rows_gloo_world2(counts):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, counts, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    expect = np.arange(sum(counts) * 5, dtype=np.float64).reshape(sum(counts), 5)
    assert np.array_equal(got[0], expect) and np.array_equal(got[1], expect)


def test_search_rows_from_counts_matches_the_explicit_problem():
    """pipeline.search_rows_from_counts (row index + T, dp2scaf left to the library) against search_problem_from_counts (explicit dp2scaf):
    scaffolds with fewer than two windows are dropped (quirk Q1, ScafDpData.cpp:92-93)."""
    from abawaca_b200 import pipeline
    rng = np.random.default_rng(5)
    for counts in (rng.integers(0, 7, 5000), rng.integers(2, 9, 300), np.array([1, 1, 5, 0, 2]), np.array([3])):
        counts = counts.astype(np.int64)
        keep, dp2scaf, T, kept = pipeline.search_problem_from_counts(counts)
        rows, T2, kept2, n = pipeline.search_rows_from_counts(counts)
        assert n == int(keep.sum()) and np.array_equal(T, T2)
        if keep.all():
            assert rows is None and kept2 is None
        else:
            assert rows.dtype == np.uint64 and np.array_equal(rows, np.nonzero(keep)[0]) and np.array_equal(kept, kept2)
        # dp2scaf as the library derives it from T: datapoints of a scaffold are consecutive
        assert np.array_equal(dp2scaf, np.repeat(np.arange(T.size, dtype=np.uint32), T))
