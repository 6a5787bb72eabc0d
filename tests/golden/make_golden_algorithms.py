"""
Golden vectors for the exported geometry helpers, produced by running the REFERENCE's own
scalar functions in the build container:

    NUMBA_CACHE_DIR=/tmp/nbcache python tests/golden/make_golden_algorithms.py

  liang_barsky_line_box_clip, cohen_sutherland_line_box_clip, cyrus_beck_line_polygon_clip  (algorithms/)
  point_in_polygon, point_in_triangle / points_in_triangles                                 (geometry_utils.py)
  close_polygons, edges, example_1d_network                                                 (demo.py; matplotlib is stubbed:
                                                                                            only the plotting helpers use it)
Inputs: coordinates drawn from a small lattice (so that endpoints on box sides and corners, zero-length segments,
axis-parallel segments on the boundary, points on vertices and edges all occur many times) mixed with continuous
random ones; the cases of the reference's own tables (tests/test_algorithms/test_line_box_clip.py:47-304: box
(0, 2, 0, 2) and the square polygon, segments through corners / along sides / touching) are all inside that lattice.
"""

import os
import pathlib
import sys
import types

os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/nbcache")
REFERENCE = os.environ.get("CELLTREE_REFERENCE", "/root/reference")
sys.path.insert(0, REFERENCE)

import numpy as np  # noqa: E402

for name in ("matplotlib", "matplotlib.tri", "matplotlib.patches", "matplotlib.collections"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules["matplotlib"].tri = sys.modules["matplotlib.tri"]
sys.modules["matplotlib"].patches = sys.modules["matplotlib.patches"]
sys.modules["matplotlib.collections"].LineCollection = object

from numba_celltree import demo as ref_demo  # noqa: E402
from numba_celltree import geometry_utils as gu  # noqa: E402
from numba_celltree.algorithms import (  # noqa: E402
    cohen_sutherland_line_box_clip,
    cyrus_beck_line_polygon_clip,
    liang_barsky_line_box_clip,
)
from numba_celltree.constants import Box, Point, Triangle  # noqa: E402

HERE = pathlib.Path(__file__).resolve().parent
LATTICE = np.array([-1.0, -0.1, 0.0, 0.5, 1.0, 1.5, 2.0, 2.1, 3.0])


def mixed_points(rng, n):
    lattice = rng.choice(LATTICE, size=(n, 2))
    smooth = rng.uniform(-1.5, 3.5, size=(n, 2))
    pick = rng.random(n) < 0.6
    return np.where(pick[:, None], lattice, smooth)


def main():
    rng = np.random.default_rng(20240611)
    out = {}
    # ---- segment / box -------------------------------------------------------------------------------------------------
    n = 6000
    a, b = mixed_points(rng, n), mixed_points(rng, n)
    same = rng.random(n) < 0.03
    b[same] = a[same]  # zero-length segments
    boxes = np.empty((n, 4))
    boxes[:] = (0.0, 2.0, 0.0, 2.0)
    odd = rng.random(n) < 0.3
    lo = rng.choice(LATTICE, size=(n, 2))
    boxes[odd, 0], boxes[odd, 2] = lo[odd, 0], lo[odd, 1]
    boxes[odd, 1] = boxes[odd, 0] + rng.choice([0.0, 0.5, 1.0, 2.0], size=odd.sum())  # also boxes of zero width
    boxes[odd, 3] = boxes[odd, 2] + rng.choice([0.0, 0.5, 1.0, 2.0], size=odd.sum())
    for name, fn in (("lb", liang_barsky_line_box_clip), ("cs", cohen_sutherland_line_box_clip)):
        hit = np.empty(n, dtype=bool)
        c = np.empty((n, 2))
        d = np.empty((n, 2))
        for i in range(n):
            hit[i], c[i], d[i] = fn(Point(*a[i]), Point(*b[i]), Box(*boxes[i]))
        out[f"{name}_hit"], out[f"{name}_c"], out[f"{name}_d"] = hit, c, d
    out["seg_a"], out["seg_b"], out["seg_boxes"] = a, b, boxes
    # ---- segment / convex polygon --------------------------------------------------------------------------------------
    polygons = {
        "square": np.array([[0.0, 0.0], [2.0, 0.0], [2.0, 2.0], [0.0, 2.0]]),
        "triangle": np.array([[0.0, 0.0], [2.0, 0.5], [0.5, 2.0]]),
        "hexagon": np.array([[1.0, -1.0], [2.5, 0.0], [2.5, 1.5], [1.0, 3.0], [-0.5, 1.5], [-0.5, 0.0]]),
    }
    m = 2500
    for pname, poly in polygons.items():
        for tol in (1e-9, 0.0):
            pa, pb = mixed_points(rng, m), mixed_points(rng, m)
            hit = np.empty(m, dtype=bool)
            c = np.empty((m, 2))
            d = np.empty((m, 2))
            for i in range(m):
                hit[i], c[i], d[i] = cyrus_beck_line_polygon_clip(Point(*pa[i]), Point(*pb[i]), poly, tol)
            key = f"cb_{pname}_{'tol' if tol else 'zero'}"
            out[f"{key}_a"], out[f"{key}_b"], out[f"{key}_poly"], out[f"{key}_tolerance"] = pa, pb, poly, np.float64(tol)
            out[f"{key}_hit"], out[f"{key}_c"], out[f"{key}_d"] = hit, c, d
    # ---- point in polygon (no tolerance) -------------------------------------------------------------------------------
    pip_polys = dict(polygons)
    pip_polys["unit_square"] = np.array([[0.0, 0.0], [1.0, 0.0], [1.0, 1.0], [0.0, 1.0]])  # tests/test_geometry_utils.py:130-146
    pip_polys["square_reversed"] = pip_polys["unit_square"][::-1].copy()
    pip_polys["concave"] = np.array([[0.0, 0.0], [3.0, 0.0], [3.0, 3.0], [1.5, 1.0], [0.0, 3.0]])
    pip_polys["repeated_vertex"] = np.array([[0.0, 0.0], [2.0, 0.0], [2.0, 0.0], [2.0, 2.0], [0.0, 2.0]])
    for pname, poly in pip_polys.items():
        pts = mixed_points(rng, 3000)
        inside = np.array([gu.point_in_polygon(Point(*p), poly) for p in pts])
        out[f"pip_{pname}_points"], out[f"pip_{pname}_poly"], out[f"pip_{pname}_inside"] = pts, poly, inside
    # ---- points in triangles -------------------------------------------------------------------------------------------
    import scipy.spatial

    verts = np.concatenate((rng.choice(LATTICE, size=(40, 2)), rng.uniform(-1, 3, size=(60, 2))))
    verts = np.unique(verts, axis=0)
    tri = scipy.spatial.Delaunay(verts).simplices.astype(np.int64)
    faces = np.full((len(tri), 4), -1, dtype=np.int64)
    faces[:, :3] = tri
    for tol in (1e-9, 1e-3, 0.0):
        pts = mixed_points(rng, 5000)
        # half of the points sit on (or next to) their triangle: combinations of its own corners
        idx = rng.integers(0, len(faces), len(pts))
        w = rng.choice([0.0, 0.25, 0.5, 1.0], size=(len(pts), 3))
        w[:, 2] = np.where(w.sum(axis=1) == 0, 1.0, w[:, 2])
        w /= w.sum(axis=1)[:, None]
        own = (verts[tri[idx]] * w[:, :, None]).sum(axis=1)
        use_own = rng.random(len(pts)) < 0.5
        pts = np.where(use_own[:, None], own, pts)
        inside = gu.points_in_triangles(pts, idx, faces, verts, tol)
        single = np.array([gu.point_in_triangle(Point(*pts[i]), Triangle(*(Point(*verts[k]) for k in tri[idx[i]])), tol) for i in range(200)])
        assert np.array_equal(single, inside[:200])
        key = f"pit_{tol:g}"
        out[f"{key}_points"], out[f"{key}_face_indices"], out[f"{key}_inside"], out[f"{key}_tolerance"] = pts, idx, inside, np.float64(tol)
    out["pit_faces"], out["pit_vertices"] = faces, verts
    # ---- demo helpers ----------------------------------------------------------------------------------------------------
    mixed = np.array([[0, 1, 4, 3], [1, 2, 4, -1], [2, 5, 4, -1], [3, 4, 7, 6], [4, 5, 8, 7]], dtype=np.int64)
    out["demo_faces"] = mixed
    out["demo_closed"] = ref_demo.close_polygons(mixed, -1)
    out["demo_edges"] = np.asarray(ref_demo.edges(mixed, -1))
    nv, ne = ref_demo.example_1d_network()
    out["demo_network_vertices"], out["demo_network_edges"] = nv, ne
    np.savez_compressed(HERE / "algorithms.npz", **out)
    print("wrote", HERE / "algorithms.npz", {k: v.shape for k, v in out.items() if k.endswith(("_hit", "_inside"))})
    print("hits:", {k: int(v.sum()) for k, v in out.items() if k.endswith(("_hit", "_inside"))})


if __name__ == "__main__":
    main()
