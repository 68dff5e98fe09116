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
