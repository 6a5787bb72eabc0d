"""Host-side sharding logic of the multi-GPU path, world_size 2 over gloo on CPU."""

import os

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from numba_celltree_b200.distributed import (
    assemble_indices,
    exchange_totals,
    gather_pairs,
    globalize_pairs,
    locate_points_sharded,
    query_pairs_sharded,
    shard_range,
)


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


class _FakeTree:
    """Stands in for a CellTree2d on a GPU-less host: deterministic fixed-size and variable-length answers per query, so
    that the sharding / offset / gather logic can be checked against the unsharded call."""

    def locate_points(self, points, tolerance=None):
        return (np.floor(points[:, 0] * 10) + 100 * np.floor(points[:, 1] * 10)).astype(np.intp)

    def intersect_boxes(self, boxes):
        counts = (np.floor(boxes[:, 0] * 7) % 4).astype(np.intp)  # 0..3 pairs per box
        i = np.repeat(np.arange(len(boxes), dtype=np.intp), counts)
        j = (np.floor(boxes[i, 1] * 1000)).astype(np.intp) + np.concatenate([np.arange(c) for c in counts] or [np.empty(0, int)])
        return i, j, boxes[i, 2] * 0.5


def _sharded_worker(rank, world, port, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        tree = _FakeTree()
        rng = np.random.default_rng(11)
        points = rng.uniform(0, 1, (1003, 2))
        boxes = rng.uniform(0, 1, (777, 4))
        lo, hi, found = locate_points_sharded(tree, points)
        np.save(os.path.join(tmp, f"points{rank}.npy"), np.concatenate([[lo, hi], found]))
        gi, j, area, offset, total, totals = query_pairs_sharded(tree, "intersect_boxes", boxes)
        assert sum(totals) == total and sum(totals[:rank]) == offset
        gathered = gather_pairs(gi, j, area, offset, total, totals, dst=0)
        if rank == 0:
            np.savez(os.path.join(tmp, "gathered.npz"), i=gathered[0], j=gathered[1], area=gathered[2])
        else:
            assert gathered is None
        with pytest.raises(ValueError):
            query_pairs_sharded(tree, "locate_faces", boxes)
    finally:
        dist.destroy_process_group()


def test_sharded_queries_reproduce_the_unsharded_order_world2(tmp_path):
    world = 2
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_sharded_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    tree = _FakeTree()
    rng = np.random.default_rng(11)
    points = rng.uniform(0, 1, (1003, 2))
    boxes = rng.uniform(0, 1, (777, 4))
    expected = tree.locate_points(points)
    out = np.full(len(points), -7, dtype=np.intp)
    for r in range(world):
        part = np.load(tmp_path / f"points{r}.npy")
        lo, hi = int(part[0]), int(part[1])
        out[lo:hi] = part[2:]
    assert np.array_equal(out, expected)
    ei, ej, ea = tree.intersect_boxes(boxes)
    got = np.load(tmp_path / "gathered.npz")
    assert np.array_equal(got["i"], ei) and np.array_equal(got["j"], ej) and np.array_equal(got["area"], ea)


def _tensor_worker(rank, world, port, tmp):
    import torch

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # tensor pieces (what the device-resident calls return; CPU tensors stand in for CUDA tensors under gloo): every
        # rank's pairs go point to point into rows [offset, offset + count) of buffers preallocated on the destination
        tree = _FakeTree()
        boxes = np.random.default_rng(4).uniform(0, 1, (501, 4))
        lo, hi = shard_range(len(boxes), rank, world)
        i, j, area = tree.intersect_boxes(boxes[lo:hi])
        offset, total, totals = exchange_totals(len(i))
        pieces = (torch.from_numpy(i + lo), torch.from_numpy(j), torch.from_numpy(area))
        got = gather_pairs(*pieces, offset, total, totals, dst=1)
        if rank == 1:
            assert all(isinstance(t, torch.Tensor) for t in got)
            np.savez(os.path.join(tmp, "tensor_gathered.npz"), i=got[0].numpy(), j=got[1].numpy(), area=got[2].numpy())
        else:
            assert got is None
        # fixed-size results of equal shards, assembled on every rank (narrowed to int32 on the way), tensors and NumPy
        points = np.random.default_rng(13).uniform(0, 1, (1000, 2))
        points[::17] = 2.0
        lo, hi, found = locate_points_sharded(tree, points)
        found = np.where(points[lo:hi, 0] > 1.0, -1, found)
        for narrow in (True, False):
            whole = assemble_indices(torch.from_numpy(found), narrow=narrow)
            assert whole.dtype == torch.int64
            np.save(os.path.join(tmp, f"assembled{rank}_{int(narrow)}.npy"), whole.numpy())
        into = torch.empty(len(points), dtype=torch.int64)
        assert assemble_indices(torch.from_numpy(found), out=into) is into
        assert np.array_equal(into.numpy(), assemble_indices(found))
    finally:
        dist.destroy_process_group()


def test_gather_pairs_moves_tensors_point_to_point_world2(tmp_path):
    world = 2
    port = 32500 + (os.getpid() % 2000)
    mp.spawn(_tensor_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    tree = _FakeTree()
    boxes = np.random.default_rng(4).uniform(0, 1, (501, 4))
    ei, ej, ea = tree.intersect_boxes(boxes)
    got = np.load(tmp_path / "tensor_gathered.npz")
    assert np.array_equal(got["i"], ei) and np.array_equal(got["j"], ej) and np.array_equal(got["area"], ea)
    points = np.random.default_rng(13).uniform(0, 1, (1000, 2))
    points[::17] = 2.0
    expected = np.where(points[:, 0] > 1.0, -1, tree.locate_points(points))
    for rank in range(world):
        for narrow in (0, 1):
            assert np.array_equal(np.load(tmp_path / f"assembled{rank}_{narrow}.npy"), expected)
