"""
Seeded synthetic meshes and query sets: the BASELINE.json configurations (C1-C5,
recipes fixed in SURVEY.md section 8d) and the small meshes the parity tests use.

Pure NumPy / SciPy host code; nothing here touches the GPU.
"""

from __future__ import annotations

import numpy as np


def quad_mesh(nx: int, ny: int):
    """Structured quads on the unit square; vertex id = j*(nx+1)+i, faces CCW, int64 (C2 / C5 recipe)."""
    x = np.linspace(0.0, 1.0, nx + 1)
    y = np.linspace(0.0, 1.0, ny + 1)
    xx, yy = np.meshgrid(x, y, indexing="xy")
    vertices = np.column_stack((xx.ravel(), yy.ravel()))
    i, j = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")
    a = (j * (nx + 1) + i).ravel()
    faces = np.column_stack((a, a + 1, a + nx + 2, a + nx + 1)).astype(np.int64)
    return vertices, faces


def delaunay_mesh(n_points: int, seed: int):
    """Delaunay triangulation of seeded uniform points on the unit square (C3 recipe: n=1_000_000, seed=1234)."""
    import scipy.spatial

    vertices = np.random.default_rng(seed).uniform(0, 1, (n_points, 2))
    faces = scipy.spatial.Delaunay(vertices).simplices.astype(np.int64)
    return vertices, faces


def generate_disk(partitions: int, depth: int):
    """
    Triangulated unit disk: the point recipe of the reference's demo.generate_disk
    (demo.py:91-123), triangulated with scipy instead of matplotlib (C1: (25, 20)).
    """
    import scipy.spatial

    if partitions < 3:
        raise ValueError("partitions should be >= 3")
    N = depth + 1
    n_per_level = partitions * np.arange(N)
    n_per_level[0] = 1
    delta_angle = (2 * np.pi) / np.repeat(n_per_level, n_per_level)
    index = np.repeat(np.insert(n_per_level.cumsum()[:-1], 0, 0), n_per_level)
    angles = delta_angle.cumsum()
    angles = angles - angles[index] + 0.5 * np.pi
    radii = np.repeat(np.linspace(0.0, 1.0, N), n_per_level)
    x = np.cos(angles) * radii
    y = np.sin(angles) * radii
    xy = np.column_stack((x, y))
    triangles = scipy.spatial.Delaunay(xy).simplices.astype(np.int64)
    return xy, triangles


def random_network(n_edges: int, seed: int):
    """A seeded 1-D network: short random segments chained into polylines (EdgeCellTree2d tests)."""
    rng = np.random.default_rng(seed)
    n_lines = max(1, n_edges // 20)
    per = n_edges // n_lines
    vertices = []
    edges = []
    for _ in range(n_lines):
        start = rng.uniform(0, 100, 2)
        steps = rng.normal(0, 1.5, (per, 2)) + rng.normal(0, 1.0, 2)
        line = np.vstack([start, start + np.cumsum(steps, axis=0)])
        base = sum(len(v) for v in vertices)
        vertices.append(line)
        idx = np.arange(per) + base
        edges.append(np.column_stack((idx, idx + 1)))
    return np.vstack(vertices), np.vstack(edges).astype(np.int64)


# --- BASELINE.json query sets (keep the draw order) -------------------------
def c1_points(n: int = 1_000_000):
    return np.random.default_rng(0).uniform(-1, 1, (n, 2))


def c2_points(n: int = 100_000_000):
    """First n of the seed-42 points (the generator streams, so a prefix equals the prefix of the full draw)."""
    return np.random.default_rng(42).uniform(0, 1, (n, 2))


def c3_boxes(n_tri: int, n: int = 10_000_000):
    h = np.sqrt(1.0 / n_tri)
    r2 = np.random.default_rng(2)
    c, wh = r2.uniform(0, 1, (n, 2)), r2.uniform(0, 4 * h, (n, 2))
    return np.column_stack((c[:, 0] - wh[:, 0] / 2, c[:, 0] + wh[:, 0] / 2, c[:, 1] - wh[:, 1] / 2, c[:, 1] + wh[:, 1] / 2))


def c4_edges(n_tri: int, n: int = 10_000_000):
    h = np.sqrt(1.0 / n_tri)
    r3 = np.random.default_rng(3)
    a0 = r3.uniform(0, 1, (n, 2))
    ang = r3.uniform(0, 2 * np.pi, n)
    L = r3.uniform(0, 10 * h, n)
    return np.stack((a0, a0 + np.column_stack((np.cos(ang), np.sin(ang))) * L[:, None]), axis=1)
