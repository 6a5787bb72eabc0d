"""Host-side sharding logic of the multi-GPU path, world_size 2 over gloo on CPU."""

import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from numba_celltree_b200.distributed import exchange_totals, globalize_pairs, shard_range


def test_shard_range_tiles_the_queries():
    for n in (0, 1, 7, 100, 101, 1_000_003):
        for world in (1, 2, 3, 8):
            ranges = [shard_range(n, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            for (a0, a1), (b0, b1) in zip(ranges[:-1], ranges[1:]):
                assert a1 == b0
            sizes = [hi - lo for lo, hi in ranges]
            assert max(sizes) - min(sizes) <= 1
            expected = [len(c) for c in np.array_split(np.arange(n), world)]
            assert sizes == expected


def _worker(rank, world, port, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # a fake variable-length query: query q yields (q % 3) pairs; each rank owns a contiguous range
        n = 1001
        lo, hi = shard_range(n, rank, world)
        counts = np.arange(lo, hi) % 3
        i_local = np.repeat(np.arange(hi - lo), counts)
        offset, total, totals = exchange_totals(len(i_local))
        i_global = globalize_pairs(i_local, lo)
        np.save(os.path.join(tmp, f"part{rank}.npy"), np.array([offset, total, *totals, *i_global]))
    finally:
        dist.destroy_process_group()


def test_offsets_allgather_world2(tmp_path):
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    n = 1001
    expected = np.repeat(np.arange(n), np.arange(n) % 3)
    parts = [np.load(tmp_path / f"part{r}.npy") for r in range(world)]
    out = np.empty(len(expected), dtype=np.int64)
    for r, part in enumerate(parts):
        offset, total = int(part[0]), int(part[1])
        totals = part[2 : 2 + world]
        body = part[2 + world :]
        assert total == len(expected)
        assert offset == totals[:r].sum()
        out[offset : offset + len(body)] = body
    assert np.array_equal(out, expected)
