"""
Multi-GPU plumbing for the query path: one process per GPU, torch.distributed (NCCL over NVLink / NVSwitch).

The path shards naturally: the tree is REPLICATED (built once on the source rank, broadcast as
device arrays), queries are split into contiguous ranges of the original query index, fixed-size results need
no collective, and variable-length results need exactly one all-gather of the per-rank pair totals to turn
local offsets into global offsets.  Concatenating the per-rank results in rank order reproduces the
single-GPU (= the reference's) output order.

The host-side logic (shard_range, exchange_totals, the *_sharded calls, assemble_indices, gather_pairs) is backend-agnostic and is
exercised with gloo on CPU in tests/test_distributed.py.
"""

from __future__ import annotations

import ctypes
from typing import Tuple

import numpy as np


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous range [lo, hi) of query indices owned by `rank` (sizes differ by at most one, np.array_split style)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def _collective_device(device=None):
    """Where the tensors of a collective must live: the given device, else the current CUDA device under NCCL, else the host."""
    import torch
    import torch.distributed as dist

    if device is not None:
        return torch.device(device)
    if dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def exchange_totals(local_total: int, device=None):
    """
    One all-gather of the per-rank result sizes -> (global offset of this rank's first pair, grand total, all totals).
    The path's only data-path collective: world_size int64 values.
    """
    import torch
    import torch.distributed as dist

    world = dist.get_world_size()
    rank = dist.get_rank()
    mine = torch.tensor([int(local_total)], dtype=torch.int64, device=_collective_device(device))
    gathered = torch.empty(world, dtype=torch.int64, device=mine.device)
    dist.all_gather_into_tensor(gathered, mine)
    totals = [int(t) for t in gathered.tolist()]
    return sum(totals[:rank]), sum(totals), totals


def globalize_pairs(i_local, lo: int):
    """Local query indices of a shard -> global query indices (the shard starts at query `lo`)."""
    return i_local + lo


def _rows(queries, lo: int, hi: int):
    """Rows [lo, hi) of a query array (ndarray, CUDA tensor or anything sliceable); lists are converted first."""
    if not hasattr(queries, "shape"):
        queries = np.asarray(queries)
    return queries[lo:hi]


def locate_points_sharded(tree, points, tolerance=None, weights: bool = False):
    """
    This rank's share of ``tree.locate_points(points)`` (or ``compute_barycentric_weights`` with ``weights=True``):
    every rank passes the SAME `points` (or at least an array of the same length whose rows [lo, hi) are valid) and
    gets ``(lo, hi, result)`` for its contiguous range.  Fixed-size results need no collective: writing each
    rank's result to rows [lo, hi) of one array reproduces the single-GPU output.
    """
    import torch.distributed as dist

    lo, hi = shard_range(len(points), dist.get_rank(), dist.get_world_size())
    mine = _rows(points, lo, hi)
    result = tree.compute_barycentric_weights(mine, tolerance) if weights else tree.locate_points(mine, tolerance)
    return lo, hi, result


def assemble_indices(local, out=None, narrow: bool = True):
    """
    Every rank's share of a fixed-size index result (``locate_points_sharded``: cell indices, -1 = outside), put together
    on EVERY rank in query order: one all-gather of equal shards.  Cell indices fit 32 bits (the tree holds fewer than
    2**31 cells, `ct_tree_info.n_elem`), so with `narrow` they travel as int32 -- half the bytes over NVLink -- and are
    widened to the API's int64 on arrival.  `local` is a torch tensor (CUDA under NCCL) or a NumPy array; all shards must
    have the same length (``n % world == 0``; otherwise keep the result sharded or use `gather_pairs`-style transfers).
    `out`: optional preallocated int64 tensor of world * len(local) entries.
    """
    import torch
    import torch.distributed as dist

    world = dist.get_world_size()
    as_numpy = isinstance(local, np.ndarray)
    piece = torch.from_numpy(np.ascontiguousarray(local)) if as_numpy else local.contiguous()
    if as_numpy and dist.get_backend() == "nccl":
        piece = piece.to(_collective_device())
    if narrow:
        piece = piece.to(torch.int32)
    gathered = torch.empty(world * piece.shape[0], dtype=piece.dtype, device=piece.device)
    dist.all_gather_into_tensor(gathered, piece)
    if out is None:
        out = gathered.to(torch.int64)
    else:
        out.copy_(gathered)
    return out.cpu().numpy() if as_numpy else out


def query_pairs_sharded(tree, method: str, queries, *args, device=None):
    """
    This rank's share of a variable-length query -- `method` is one of ``locate_boxes``, ``intersect_boxes``,
    ``intersect_edges`` (queries = boxes / segments) -- over the contiguous range of `queries` this rank owns.
    Returns ``(i_global, j, payload_or_None, offset, total, totals)``: the local pairs with GLOBAL query indices, the
    position of this rank's first pair in the concatenated (= single-GPU = reference) result, the grand total and every
    rank's count, obtained with the path's one collective (an all-gather of the per-rank pair counts).
    """
    import torch.distributed as dist

    if method not in ("locate_boxes", "intersect_boxes", "intersect_edges"):
        raise ValueError(f"query_pairs_sharded: unsupported method {method!r}")
    lo, hi = shard_range(len(queries), dist.get_rank(), dist.get_world_size())
    out = getattr(tree, method)(_rows(queries, lo, hi), *args)
    i_local, j = out[0], out[1]
    payload = out[2] if len(out) > 2 else None
    offset, total, totals = exchange_totals(len(i_local), device=device)
    return globalize_pairs(i_local, lo), j, payload, offset, total, totals


def intersect_faces_sharded(tree, vertices, faces, fill_value: int, device=None):
    """
    This rank's share of ``tree.intersect_faces(vertices, faces, fill_value)``: the query FACES are split into contiguous
    ranges (every rank holds the whole vertex array, which its faces index into); same return value as
    `query_pairs_sharded`.  Regridding: each rank ends up with the weight triplets of its own target cells.
    """
    import torch.distributed as dist

    lo, hi = shard_range(len(faces), dist.get_rank(), dist.get_world_size())
    i_local, j, area = tree.intersect_faces(vertices, _rows(faces, lo, hi), fill_value)
    offset, total, totals = exchange_totals(len(i_local), device=device)
    return globalize_pairs(i_local, lo), j, area, offset, total, totals


def gather_pairs(i_global, j, payload, offset: int, total: int, totals=None, dst: int = 0):
    """
    Assemble the sharded pairs on rank `dst` in the reference's order (other ranks get None).  Every rank's piece goes
    straight to its place -- rows [offset, offset + count) of buffers preallocated on `dst` -- by point-to-point transfers:
    CUDA tensors travel over NCCL (NVLink, device to device, nothing touches the host) and the result stays on `dst`'s
    GPU; NumPy pieces travel as host tensors over whatever backend the group has and come back as NumPy arrays.
    """
    import torch
    import torch.distributed as dist

    rank, world = dist.get_rank(), dist.get_world_size()
    if totals is None:
        _, _, totals = exchange_totals(len(i_global))
    on_device = not isinstance(i_global, np.ndarray)

    def as_tensor(a):
        if a is None:
            return None
        return torch.from_numpy(np.ascontiguousarray(a)) if isinstance(a, np.ndarray) else a.contiguous()

    pieces = [as_tensor(i_global), as_tensor(j), as_tensor(payload)]
    if not on_device and dist.get_backend() == "nccl":
        pieces = [None if p is None else p.to(_collective_device()) for p in pieces]
    if rank != dst:
        for piece in pieces:
            if piece is not None and piece.shape[0] > 0:
                dist.send(piece, dst)
        return None
    starts = np.concatenate(([0], np.cumsum(totals)))
    out = []
    for piece in pieces:
        if piece is None:
            out.append(None)
            continue
        whole = torch.empty((int(total),) + tuple(piece.shape[1:]), dtype=piece.dtype, device=piece.device)
        for r in range(world):
            lo, hi = int(starts[r]), int(starts[r + 1])
            if hi == lo:
                continue
            if r == dst:
                whole[lo:hi] = piece
            else:
                dist.recv(whole[lo:hi], r)
        out.append(whole)
    if on_device:
        return tuple(out)
    return tuple(None if t is None else t.cpu().numpy() for t in out)


def export_device_arrays(tree, device):
    """Device copies (torch CUDA tensors) of everything ct_tree_from_arrays needs, in the C-ABI's layouts."""
    import torch

    from numba_celltree_b200 import _lib

    info = tree._tree.info
    n, m, n_nodes = int(info.n_elem), int(info.n_max_vert), int(info.n_nodes)
    arrays = {
        "vertices": torch.from_numpy(tree.vertices).to(device),
        "elements": torch.empty((n, m), dtype=torch.int64, device=device),
        "nodes": torch.empty(n_nodes * 41, dtype=torch.uint8, device=device),
        "bb_indices": torch.empty(n, dtype=torch.int64, device=device),
        "bb_coords": torch.empty((n, 4), dtype=torch.float64, device=device),
    }
    _lib.check(
        _lib.load().ct_tree_download(
            tree._tree.handle, arrays["nodes"].data_ptr(), arrays["bb_indices"].data_ptr(), arrays["bb_coords"].data_ptr(),
            arrays["elements"].data_ptr(), None, _lib.CT_MEM_DEVICE,
        )
    )  # fmt: skip
    torch.cuda.synchronize(device)
    meta = {
        "n_vertex": int(info.n_vertex), "n_elem": n, "n_max_vert": m, "n_nodes": n_nodes, "kind": int(info.kind),
        "n_buckets": int(info.n_buckets), "cells_per_leaf": int(info.cells_per_leaf), "cls": type(tree).__name__,
    }  # fmt: skip
    return meta, arrays


def import_device_arrays(meta, arrays):
    """Build a tree object around device arrays received from another rank (no build kernels run)."""
    from numba_celltree_b200 import CellTree2d, EdgeCellTree2d, _lib
    from numba_celltree_b200.celltree_base import DeviceTree

    cls = CellTree2d if meta["cls"] == "CellTree2d" else EdgeCellTree2d
    handle = ctypes.c_void_p()
    _lib.check(
        _lib.load().ct_tree_from_arrays(
            arrays["vertices"].data_ptr(), meta["n_vertex"], arrays["elements"].data_ptr(), meta["n_elem"], meta["n_max_vert"],
            meta["kind"], arrays["nodes"].data_ptr(), meta["n_nodes"], arrays["bb_indices"].data_ptr(),
            arrays["bb_coords"].data_ptr(), meta["cells_per_leaf"], _lib.CT_MEM_DEVICE, ctypes.byref(handle),
        )
    )  # fmt: skip
    tree = cls.__new__(cls)
    tree._tree = DeviceTree(handle.value)
    tree.vertices = arrays["vertices"].cpu().numpy()
    tree.n_buckets = meta["n_buckets"]
    tree.cells_per_leaf = meta["cells_per_leaf"]
    if cls is EdgeCellTree2d:
        tree.edges = arrays["elements"].cpu().numpy()
    return tree


def broadcast_tree(tree, src: int, device):
    """Replicate the tree of rank `src` on every rank: one broadcast per device array (NCCL over NVLink)."""
    import torch
    import torch.distributed as dist

    rank = dist.get_rank()
    if rank == src:
        meta, arrays = export_device_arrays(tree, device)
        box = [meta]
    else:
        box = [None]
    dist.broadcast_object_list(box, src=src)
    meta = box[0]
    if rank != src:
        n, m = meta["n_elem"], meta["n_max_vert"]
        arrays = {
            "vertices": torch.empty((meta["n_vertex"], 2), dtype=torch.float64, device=device),
            "elements": torch.empty((n, m), dtype=torch.int64, device=device),
            "nodes": torch.empty(meta["n_nodes"] * 41, dtype=torch.uint8, device=device),
            "bb_indices": torch.empty(n, dtype=torch.int64, device=device),
            "bb_coords": torch.empty((n, 4), dtype=torch.float64, device=device),
        }
    for key in ("vertices", "elements", "nodes", "bb_indices", "bb_coords"):
        dist.broadcast(arrays[key], src=src)
    torch.cuda.synchronize(device)
    if rank == src:
        return tree
    return import_device_arrays(meta, arrays)
