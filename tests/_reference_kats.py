"""
Known-answer tests restating the reference's own API-level tests (tests/test_celltree.py and
tests/test_edgecelltree.py of Deltares/numba_celltree): same inputs, same expected values, parameterised by
the implementation under test (the CPU oracle or the CUDA package).  file:line cites are into the reference.
"""

import numpy as np
import pytest

from numba_celltree_b200.synthetic import generate_disk

# two triangles -- tests/test_celltree.py:20-33
NODES2 = [[0.0, 0.0], [2.0, 0.0], [1.0, 2.0], [3.0, 2.0]]
FACES2 = [[0, 1, 2], [1, 3, 2]]

# 21 triangles -- tests/test_celltree.py:36-82
NODES21 = [(5, 1), (10, 1), (3, 3), (7, 3), (9, 4), (12, 4), (5, 5), (3, 7), (5, 7), (7, 7), (9, 7), (11, 7), (5, 9), (8, 9),
           (11, 9), (9, 11), (11, 11), (7, 13), (9, 13), (7, 15)]  # fmt: skip
FACES21 = [(0, 1, 3), (0, 2, 6), (0, 3, 6), (1, 3, 4), (1, 4, 5), (2, 6, 7), (6, 7, 8), (7, 8, 12), (6, 8, 9), (8, 9, 12),
           (9, 12, 13), (4, 5, 11), (4, 10, 11), (9, 10, 13), (10, 11, 14), (10, 13, 14), (13, 14, 15), (14, 15, 16),
           (15, 16, 18), (15, 17, 18), (17, 18, 19)]  # fmt: skip

# 12-vertex block -- tests/test_celltree.py:200-216
NODES12 = np.array(
    [[0.0, 0.0], [0.0, 2.0], [2.0, 0.0], [2.0, 2.0], [4.0, 0.0], [4.0, 2.0], [6.0, 0.0], [6.0, 2.0], [0.0, 4.0], [2.0, 4.0],
     [4.0, 4.0], [6.0, 4.0]]
)  # fmt: skip
MIXED = np.array([[0, 8, 9, 5, 2], [9, 11, 7, 5, -1], [4, 7, 6, -1, -1]], dtype=np.intc)


def check_init_and_casting(CellTree2d):
    nodes = np.array(NODES2, dtype=np.float64)
    faces = np.array(FACES2, dtype=np.intc)
    CellTree2d(nodes, faces, -1)  # :95-105
    CellTree2d(nodes, faces, -1, n_buckets=2, cells_per_leaf=1)
    CellTree2d(nodes, faces, -1, n_buckets=4)
    CellTree2d(nodes, faces, -1, cells_per_leaf=2)
    CellTree2d(NODES2, FACES2, -1)  # lists, :115-117
    CellTree2d(np.array(NODES2, dtype=np.float32), np.array(FACES2, dtype=np.int32), -1)  # :120-124
    tree = CellTree2d(nodes, np.array([[0, 1, 2, -999], [1, 3, 2, -999]]), -999)  # :127-131
    assert tree.faces[0, -1] == -1
    assert tree.faces[1, -1] == -1


def check_errors(CellTree2d):
    faces = [0, 1, 2, 1, 3, 2]  # :134-156
    nodes = [(1, 2, 3), (3, 4, 5), (4, 5, 6)]
    box_coords = np.array([0.0, 1.0, 2.0, 3.0])
    edge_coords = np.array([[0.0, 1.0], [2.0, 3.0]])
    with pytest.raises(ValueError):
        CellTree2d(nodes, FACES2, -1)
    with pytest.raises(ValueError):
        CellTree2d(NODES2, faces, -1)
    tree = CellTree2d(NODES2, FACES2, -1)
    with pytest.raises(ValueError):
        tree.locate_points(nodes)
    with pytest.raises(ValueError):
        tree.intersect_faces(nodes, FACES2, -1)
    with pytest.raises(ValueError):
        tree.intersect_faces(faces, NODES2, -1)
    with pytest.raises(ValueError):
        tree.locate_boxes(box_coords)
    with pytest.raises(ValueError):
        tree.intersect_boxes(box_coords)
    with pytest.raises(ValueError):
        tree.intersect_edges(edge_coords)
    with pytest.raises(ValueError):  # :159-164
        CellTree2d(NODES2, FACES2, -1, cells_per_leaf=-1)
    with pytest.raises(ValueError):
        CellTree2d(NODES2, FACES2, -1, n_buckets=0)


def check_point_lookups(CellTree2d):
    tree = CellTree2d(np.array(NODES2), np.array(FACES2, dtype=np.intc), -1)
    assert np.array_equal(tree.locate_points(np.array([[1.0, 1.0], [2.0, 1.0], [-1.0, 1.0]])), [0, 1, -1])  # :167-178
    point = np.array([[-0.09, 0.0], [2.0, 1.0], [-1.0, 1.0]])  # :181-195
    assert np.array_equal(tree.locate_points(point, tolerance=1e-9), [-1, 1, -1])
    assert np.array_equal(tree.locate_points(point, tolerance=1e-1), [0, 1, -1])

    quads = np.array([[0, 2, 3, 1], [4, 6, 7, 5]], dtype=np.intc)  # :198-256
    pentas = np.array([[0, 8, 9, 5, 2], [9, 11, 6, 2, 5]], dtype=np.intc)
    tree1 = CellTree2d(NODES12, quads, -1, n_buckets=2, cells_per_leaf=1)
    assert np.array_equal(tree1.locate_points(np.array([[1.0, 1.0], [5.0, 1.0], [-1.0, 1.0]])), [0, 1, -1])
    tree2 = CellTree2d(NODES12, pentas, -1, n_buckets=2, cells_per_leaf=1)
    assert np.array_equal(tree2.locate_points(np.array([[1.0, 2.0], [5.0, 2.0], [-1.0, 2.0]])), [0, 1, -1])

    tree = CellTree2d(NODES12, MIXED, -1, n_buckets=2, cells_per_leaf=1)  # :259-313
    assert np.array_equal(tree.locate_points(np.array([[1.0, 1.0], [5.0, 0.5], [5.0, 3.0], [-1.0, 1.0]])), [0, 2, 1, -1])
    point = np.array(
        [[-9e-9, 0.0], [2.0, -9e-9], [-9e-9, 1.0], [1.0, -9e-9], [-1.1e-8, 1.0], [1.0, -1.1e-8], [-1.1e-8, 0.0], [2.0, -1.5e-8]]
    )
    assert np.array_equal(tree.locate_points(point), [-1] * 8)
    assert np.array_equal(tree.locate_points(point, tolerance=1e-8), [0, 0, 0, 0, -1, -1, -1, -1])

    tree = CellTree2d(NODES21, FACES21, -1)  # :316-330
    points = [(4.2, 3.0), (7.7, 13.5), (3.4, 7.000000001), (7.0, 5.0), (8.66, 10.99), (7.3, 0.74), (2.5, 5.5), (9.8, 12.3)]
    assert np.array_equal(tree.locate_points(points), (1, 20, 7, -1, -1, -1, -1, -1))


def check_box_and_edge_lookup(CellTree2d):
    tree = CellTree2d(NODES12, MIXED, -1, n_buckets=2, cells_per_leaf=1)  # :333-378
    box_coords = np.array(
        [[1.0, 2.0, 1.0, 2.0], [4.0, 5.0, 0.0, 1.0], [4.0, 5.0, 2.0, 3.0], [-1.0, 0.0, 0.0, 4.0], [6.0, 8.0, 0.0, 4.0],
         [0.0, 6.0, -1.0, 0.0], [0.0, 6.0, 4.0, 5.0]]
    )  # fmt: skip
    i, j = tree.locate_boxes(box_coords)
    assert np.array_equal(i, [0, 1, 2]) and np.array_equal(j, [0, 2, 1])
    i, j, _ = tree.intersect_boxes(box_coords)
    assert np.array_equal(i, [0, 1, 2]) and np.array_equal(j, [0, 2, 1])

    edge_coords = np.array(
        [[[1.0, 1.0], [2.0, 2.0]], [[4.0, 3.0], [5.0, 4.0]], [[5.0, 0.0], [6.0, 1.0]], [[-2.0, -1.0], [0.0, 1.0]],
         [[-2.0, -1.0], [-2.0, -1.0]]]
    )  # fmt: skip  :381-420
    i, j, xy = tree.intersect_edges(edge_coords)
    assert np.array_equal(i, [0, 1, 2]) and np.array_equal(j, [0, 1, 2])
    assert np.allclose(xy, edge_coords[:3])
    i, j, xy = tree.intersect_edges(edge_coords[:, ::-1])
    assert np.array_equal(i, [0, 1, 2]) and np.array_equal(j, [0, 1, 2])
    assert np.allclose(xy, edge_coords[:3][:, ::-1])


def disk():
    vertices, faces = generate_disk(5, 5)  # tests/test_celltree.py:85-92
    centroids = vertices[faces].mean(axis=1)
    rsquared = (centroids[:, 0] - -1.0) ** 2 + (centroids[:, 1] - -1.0) ** 2
    order = np.argsort(rsquared)
    return vertices, faces[order]


def check_example_material(CellTree2d):
    """tests/test_celltree.py:423-525 -- the exact output ORDER of every query type is part of the contract."""
    vertices, faces = disk()
    vertices += 1.0
    vertices *= 5.0
    tree = CellTree2d(vertices, faces, -1)
    assert np.array_equal(tree.locate_points(np.array([[-5.0, 1.0], [4.5, 2.5], [6.5, 4.5]])), [-1, 24, 63])

    box_coords = np.array([[4.0, 8.0, 4.0, 6.0], [0.0, 8.0, 8.0, 10.0], [10.0, 13.0, 2.0, 8.0]])
    expected_i = [0] * 30 + [1] * 23
    expected_j = [107, 109, 99, 106, 95, 93, 92, 80, 62, 73, 86, 97, 81, 79, 65, 63, 76, 57, 44, 59, 54, 58, 50, 45, 43, 31, 41,
                  36, 38, 27, 123, 112, 120, 117, 121, 108, 110, 115, 116, 101, 104, 103, 111, 88, 98, 90, 100, 72, 84, 85, 66,
                  75, 69]  # fmt: skip
    i, j = tree.locate_boxes(box_coords)
    assert np.array_equal(i, expected_i) and np.array_equal(j, expected_j)

    triangle_vertices = np.array([[5.0, 3.0], [7.0, 3.0], [7.0, 5.0], [0.0, 6.0], [4.0, 4.0], [6.0, 10.0]])
    triangles = np.array([[0, 1, 2], [3, 4, 5]])
    expected_i = [0] * 10 + [1] * 35
    expected_j = [81, 79, 61, 65, 63, 76, 44, 59, 48, 40, 115, 116, 101, 104, 91, 103, 111, 88, 98, 83, 72, 84, 68, 74, 54, 58,
                  55, 67, 47, 35, 50, 66, 46, 53, 26, 34, 49, 56, 29, 37, 43, 36, 38, 27, 30]  # fmt: skip
    i, j, _ = tree.intersect_faces(triangle_vertices, triangles, -1)
    assert np.array_equal(i, expected_i) and np.array_equal(j, expected_j)

    edge_coords = np.array([[[0.0, 0.0], [10.0, 10.0]], [[0.0, 10.0], [10.0, 0.0]]])
    expected_i = [0] * 18 + [1] * 18
    expected_j = [0, 2, 6, 11, 16, 21, 27, 36, 43, 62, 73, 80, 93, 96, 108, 112, 120, 124, 69, 75, 66, 72, 67, 55, 50, 58, 54,
                  57, 63, 59, 65, 61, 70, 64, 77, 71]  # fmt: skip
    i, j, _ = tree.intersect_edges(edge_coords)
    assert np.array_equal(i, expected_i) and np.array_equal(j, expected_j)


def check_barycentric(CellTree2d):
    tree = CellTree2d(np.array(NODES2), np.array(FACES2, dtype=np.intc), -1)  # :528-548
    idx, w = tree.compute_barycentric_weights(np.array([[0.0, 0.0], [1.0, 1.0], [2.0, 1.0]]))
    assert np.array_equal(idx, [0, 0, 1])
    assert np.allclose(w, [[1.0, 0.0, 0.0], [0.25, 0.25, 0.5], [0.5, 0.25, 0.25]])
    nodes = np.array([[0.0, 0.0], [2.0, 0.0], [2.0, 2.0], [0.0, 2.0], [4.0, 0.0], [4.0, 4.0]])  # :551-588
    faces = np.array([[0, 1, 2, 3], [1, 4, 5, 2]])
    tree = CellTree2d(nodes, faces, -1)
    idx, w = tree.compute_barycentric_weights(np.array([[0.0, 0.0], [1.0, 1.0], [2.0, 1.0]]))
    assert np.array_equal(idx, [0, 0, 0])
    assert np.allclose(w, [[1.0, 0.0, 0.0, 0.0], [0.25, 0.25, 0.25, 0.25], [0.0, 0.5, 0.5, 0.0]])
    idx, w = tree.compute_barycentric_weights(np.array([[-3.0, 0.0]]))
    assert np.array_equal(idx, [-1]) and np.array_equal(w, np.zeros((1, 4)))


def check_locate_point_on_edge(CellTree2d):
    nodes = np.array([[0.0, 0.0], [3.0, 0.0], [1.0, 1.0], [0.0, 2.0], [3.0, 2.0]])  # :642-710
    faces = np.array([[0, 1, 2], [0, 2, 3], [2, 4, 3]])
    tree = CellTree2d(nodes, faces, -1, n_buckets=4)
    d = np.array([0.0, 0.01, 0.05, 0.15, 0.25, 0.35, 0.45, 0.55, 0.65, 0.75])
    assert (tree.locate_points(np.column_stack((d, d))) != -1).all()
    nodes = np.array([[0.0, 0.0], [2.0, 0.0], [2.0, 2.0], [0.0, 2.0], [4.0, 0.0], [4.0, 4.0], [0.0, 4.0]])
    faces = np.array([[0, 1, 2, 3], [1, 4, 5, 2], [3, 2, 5, 6]])
    tree = CellTree2d(nodes, faces, -1, n_buckets=4)
    points = np.array([[-1e-9, 1.0], [4.0 + 1e-9, 1.0], [2.0, 2.0], [3.0, -1e-9], [3.0, 4.0 + 1e-9]])
    np.testing.assert_array_equal(tree.locate_points(points, tolerance=0.9e-9), [-1, -1, 2, -1, -1])
    assert (tree.locate_points(points, tolerance=1.1e-9) != -1).all()


def check_diagnostics(CellTree2d, golden):
    g = golden("voronoi_74")  # tests/test_celltree.py:591-639 on the reference's Voronoi mesh
    tree = CellTree2d(g["vertices"], g["faces"], -1)
    bounds = tree.node_bounds
    xmin, xmax, ymin, ymax = tree.bbox
    assert bounds.shape == (len(tree.celltree_data.nodes), 4)
    assert (bounds[:, 0] >= xmin).all() and (bounds[:, 1] <= xmax).all()
    assert (bounds[:, 2] >= ymin).all() and (bounds[:, 3] <= ymax).all()
    assert tree.validate_node_bounds().all()
    tree.nodes[71]["Lmax"] = -0.02319655
    assert not tree.validate_node_bounds()[73]
    g = golden("triangles_538")
    tree = CellTree2d(g["vertices"], g["faces"], -1, n_buckets=4)
    centroids = g["vertices"][g["faces"]].mean(axis=1)
    assert np.array_equal(tree.locate_points(centroids), np.arange(len(centroids)))  # :621-628
    d = tree.to_dict_of_lists()
    assert isinstance(d, dict)
    assert list(d.keys()) == list(range(len(tree.celltree_data.nodes)))
    assert max(len(v) for v in d.values()) == 2


def check_edge_tree(EdgeCellTree2d, CellTreeData):
    vertices = np.array([[0.0, 0.0], [1.0, 0.0], [2.0, 0.0], [2.0, 1.0]], dtype=float)  # tests/test_edgecelltree.py:8-39
    edges = np.array([[0, 1], [1, 2], [2, 3]], dtype=np.int32)
    tree = EdgeCellTree2d(vertices, edges)
    assert tree.vertices.shape == (4, 2) and tree.edges.shape == (3, 2)
    assert tree.n_buckets == 4 and tree.cells_per_leaf == 2
    assert tree.nodes.shape == (3,) and tree.bb_indices.shape == (3,)
    assert tree.bb_coords.shape == (3, 4) and tree.bbox.shape == (4,)
    assert isinstance(tree.celltree_data, CellTreeData)
    np.testing.assert_array_equal(tree.bb_indices, [0, 1, 2])
    np.testing.assert_allclose(tree.bb_coords, [[0.0, 1.0, 0.0, 0.0], [1.0, 2.0, 0.0, 0.0], [2.0, 2.0, 0.0, 1.0]], atol=1e-9)
    np.testing.assert_allclose(tree.bbox, [0.0, 2.0, 0.0, 1.0], atol=1e-9)
    np.testing.assert_array_equal(tree.locate_points(np.array([[0.5, 0.0], [1.5, 0.0], [2.0, 0.5]])), [0, 1, 2])  # :42-53
    np.testing.assert_array_equal(tree.locate_points(np.array([[0.5, 0.5], [1.5, 0.5], [2.0, 0.5]])), [-1, -1, 2])
    big = np.array([[171805.657000002, 563516.366], [171889.594000001, 563437.333000001]])  # :56-69
    tree_big = EdgeCellTree2d(big, np.array([[0, 1]]))
    points = np.array([[171882.49385095935, 563444.0183244612]])
    np.testing.assert_array_equal(tree_big.locate_points(points, tolerance=1e-8), [0])
    np.testing.assert_array_equal(tree_big.locate_points(points, tolerance=1e-9), [0])
    edge_coords = np.array(
        [[[1.0, -1.0], [1.0, 1.0]], [[3.0, 1.0], [-1.0, -1.0]], [[0.0, -1.0], [0.0, 1.0]], [[-2.0, -1.0], [-3.0, -1.0]]]
    )  # :72-92
    i, j, xy = tree.intersect_edges(edge_coords)
    np.testing.assert_array_equal(i, [0, 1, 1])
    np.testing.assert_array_equal(j, [0, 2, 1])
    np.testing.assert_allclose(xy, [[1.0, 0.0], [2.0, 0.5], [1.0, 0.0]], atol=1e-9)
