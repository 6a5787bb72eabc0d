"""
The sharded (one process per GPU) query path against the single-GPU result: tree built on rank 0 and replicated with
NCCL broadcasts, queries split into contiguous ranges, one all-gather of the pair totals.
Needs two GPUs; skipped on a single-GPU box (the host-side logic runs over gloo in tests/test_distributed.py).
"""

import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, tmp):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    device = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=device)
    try:
        from numba_celltree_b200 import CellTree2d, _lib
        from numba_celltree_b200 import distributed as ctd
        from numba_celltree_b200.synthetic import c3_boxes, c4_edges, delaunay_mesh, quad_mesh

        _lib.check(_lib.load().ct_set_device(rank))
        vertices, faces = delaunay_mesh(60_000, seed=77)
        tree = CellTree2d(vertices, faces, -1) if rank == 0 else None
        tree = ctd.broadcast_tree(tree, src=0, device=device)
        rng = np.random.default_rng(5)
        points = rng.uniform(-0.02, 1.02, (400_001, 2))
        boxes = c3_boxes(len(faces), 50_001)
        edges = c4_edges(len(faces), 20_001)
        qv, qf = quad_mesh(90, 70)

        lo, hi, (found, weights) = ctd.locate_points_sharded(tree, points, weights=True)
        # device-resident shard of the same queries
        d_lo, d_hi, d_found = ctd.locate_points_sharded(tree, torch.from_numpy(points).to(device))
        assert (d_lo, d_hi) == (lo, hi) and np.array_equal(d_found.cpu().numpy(), found)
        np.savez(os.path.join(tmp, f"points{rank}.npz"), lo=lo, hi=hi, found=found, weights=weights)

        gathered = {}
        for name, queries in (("locate_boxes", boxes), ("intersect_boxes", boxes), ("intersect_edges", edges)):
            pieces = ctd.query_pairs_sharded(tree, name, queries, device=device)
            gathered[name] = ctd.gather_pairs(*pieces, dst=0)
        pieces = ctd.intersect_faces_sharded(tree, qv, qf, -1, device=device)
        gathered["intersect_faces"] = ctd.gather_pairs(*pieces, dst=0)

        if rank == 0:
            # rank 0 also answers everything alone: the concatenated shards must be identical, order included
            whole_found, whole_weights = tree.compute_barycentric_weights(points)
            np.savez(os.path.join(tmp, "whole_points.npz"), found=whole_found, weights=whole_weights)
            report = {}
            for name, queries in (("locate_boxes", boxes), ("intersect_boxes", boxes), ("intersect_edges", edges)):
                whole = getattr(tree, name)(queries)
                got = gathered[name]
                ok = np.array_equal(got[0], whole[0]) and np.array_equal(got[1], whole[1])
                if len(whole) > 2:
                    ok = ok and np.array_equal(got[2], whole[2], equal_nan=True)
                report[name] = bool(ok) and len(whole[0]) > 0
            whole = tree.intersect_faces(qv, qf, -1)
            got = gathered["intersect_faces"]
            report["intersect_faces"] = bool(
                np.array_equal(got[0], whole[0]) and np.array_equal(got[1], whole[1]) and np.array_equal(got[2], whole[2]) and len(whole[0]) > 0
            )
            np.save(os.path.join(tmp, "report.npy"), np.array([report[k] for k in sorted(report)]))
        else:
            assert all(v is None for v in gathered.values())
    finally:
        dist.destroy_process_group()


def test_two_gpus_reproduce_the_single_gpu_results(tmp_path):
    torch = pytest.importorskip("torch")
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp

    world = 2
    port = 33500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    whole = np.load(tmp_path / "whole_points.npz")
    found = np.full(len(whole["found"]), -9, dtype=np.intp)
    weights = np.full(whole["weights"].shape, np.nan)
    for r in range(world):
        part = np.load(tmp_path / f"points{r}.npz")
        found[int(part["lo"]) : int(part["hi"])] = part["found"]
        weights[int(part["lo"]) : int(part["hi"])] = part["weights"]
    assert np.array_equal(found, whole["found"])
    assert np.array_equal(weights, whole["weights"])
    assert (found >= 0).any() and (found == -1).any()
    assert np.load(tmp_path / "report.npy").all()
