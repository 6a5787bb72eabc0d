"""
BASELINE.json's full-size configurations on the GPU, checked against fingerprints of the REFERENCE's own
output at that scale (sha256 over `.tobytes()`, first 16 hex digits, recorded in SURVEY.md section 8 from runs
of Deltares/numba_celltree v0.4.2 under Numba) and through size-independent properties.
"""

import hashlib

import numpy as np
import pytest

import oracle
from numba_celltree_b200.synthetic import c2_points, c3_boxes, c4_edges, delaunay_mesh, quad_mesh

pytestmark = pytest.mark.gpu


def fingerprint(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


@pytest.fixture(scope="module")
def pkg():
    import numba_celltree_b200

    return numba_celltree_b200


def test_c2_tree_and_points_match_reference_fingerprints(pkg):
    vertices, faces = quad_mesh(4096, 4096)
    tree = pkg.CellTree2d(vertices, faces, -1)
    assert tree.nodes.shape == (16_777_215,)
    assert fingerprint(tree.nodes) == "b0f8c0fbda4936c2"
    assert fingerprint(tree.bb_indices) == "5c91b305876c6679"
    assert tree.depth == 24
    assert tree._default_tolerance() == 1e-15
    points = c2_points(10_000_000)
    result = tree.locate_points(points)
    assert fingerprint(result) == "3804707094eec09c"
    # structured mesh: the cell of a point is known in closed form away from grid lines
    ij = np.floor(points * 4096).astype(np.int64)
    interior = (np.abs(points * 4096 - np.round(points * 4096)) > 1e-6).all(axis=1)
    assert np.array_equal(result[interior], (ij[:, 1] * 4096 + ij[:, 0])[interior])
    # Wachspress weights: partition of unity and reproduction of the point (size-independent properties)
    idx, w = tree.compute_barycentric_weights(points[:2_000_000])
    assert np.array_equal(idx, result[:2_000_000])
    np.testing.assert_allclose(w.sum(axis=1), 1.0, rtol=0, atol=1e-9)
    xy = np.einsum("nk,nkd->nd", w, vertices[faces[idx]])
    np.testing.assert_allclose(xy, points[:2_000_000], rtol=0, atol=1e-9)


def test_c3_tree_matches_reference_fingerprints(pkg):
    vertices, faces = delaunay_mesh(1_000_000, seed=1234)
    assert len(faces) == 1_999_960
    tree = pkg.CellTree2d(vertices, faces, -1)
    assert tree.nodes.shape == (2_506_771,)
    assert fingerprint(tree.faces) == "14af71298a42fbde"
    assert fingerprint(tree.nodes) == "180bb5ef6edd223f"
    assert fingerprint(tree.bb_indices) == "33a963de8e7823b2"
    # C5: regridding overlap, the reference's full-scale pair count and total area
    qv, qf = quad_mesh(1000, 1000)
    i, j, area = tree.intersect_faces(qv, qf, -1)
    assert len(i) == 7_348_217
    assert abs(area.sum() - 0.999965944659) < 1e-11
    assert np.all(np.diff(i) >= 0)


def test_c3_boxes_and_c4_edges_at_full_size(pkg):
    """10 M boxes / 10 M segments against the 2 M-triangle tree: totals, ordering, and the first 150 000 queries'
    pairs (which do not depend on the rest of the batch) bit for bit against the oracle."""
    vertices, faces = delaunay_mesh(1_000_000, seed=1234)
    tree = pkg.CellTree2d(vertices, faces, -1)
    ref = oracle.CellTree2d(vertices, faces, -1)
    head = 150_000

    boxes = c3_boxes(len(faces), 10_000_000)
    i, j, area = tree.intersect_boxes(boxes)
    assert len(i) == 111_246_104 and np.all(np.diff(i) >= 0) and (area > 0).all()
    ri, rj, ra = ref.intersect_boxes(boxes[:head])
    k = np.searchsorted(i, head)
    assert np.array_equal(i[:k], ri) and np.array_equal(j[:k], rj) and np.array_equal(area[:k], ra)
    del i, j, area
    i, j = tree.locate_boxes(boxes)
    assert len(i) == 126_457_665 and np.all(np.diff(i) >= 0)
    ri, rj = ref.locate_boxes(boxes[:head])
    k = np.searchsorted(i, head)
    assert np.array_equal(i[:k], ri) and np.array_equal(j[:k], rj)
    del i, j, boxes

    edges = c4_edges(len(faces), 10_000_000)
    i, j, xy = tree.intersect_edges(edges)
    assert len(i) == 86_544_887 and np.all(np.diff(i) >= 0)
    ri, rj, rxy = ref.intersect_edges(edges[:head])
    k = np.searchsorted(i, head)
    assert np.array_equal(i[:k], ri) and np.array_equal(j[:k], rj) and np.array_equal(xy[:k], rxy, equal_nan=True)
    # along every segment the pieces come in order of t = (c - a) . (b - a) (sort_intersections_by_edge)
    a, b = edges[i, 0], edges[i, 1]
    t = ((xy[:, 0] - a) * (b - a)).sum(axis=1)
    same = i[1:] == i[:-1]
    assert not np.any(same & (t[1:] < t[:-1]))
