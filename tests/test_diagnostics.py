"""node_bounds / validate_node_bounds / to_dict_of_lists: the level-wise NumPy versions against a literal
restatement of the reference's stack loops (query.py:568-664, celltree_base.py:80-114), on trees built by the oracle."""

import numpy as np
import pytest

import oracle
from numba_celltree_b200.celltree_base import CellTree2dBase
from numba_celltree_b200.synthetic import delaunay_mesh, quad_mesh


class MirrorsOnly(CellTree2dBase):
    """The diagnostics only read the host mirrors: stand-in without a device tree."""

    nodes = bb_indices = bb_coords = bbox = None

    def __init__(self, ref):
        data = ref.celltree_data
        self.nodes, self.bb_indices, self.bb_coords, self.bbox = data.nodes.copy(), data.bb_indices, data.bb_coords, data.bbox

    def _elements(self):
        return None

    def locate_points(self, points, tolerance=None):
        raise NotImplementedError


def reference_bounds(nodes, bbox):  # collect_node_bounds, query.py:568-621
    bounds = np.empty((len(nodes), 4))
    bounds[0] = bbox
    stack = [(2, 0, 0), (1, 0, 1)] if nodes[0]["child"] != -1 else []
    while stack:
        index, parent, side = stack.pop()
        bounds[index] = bounds[parent]
        dim = 1 if nodes[parent]["dim"] else 0
        bounds[index, 2 * dim + side] = nodes[parent]["Lmax"] if side else nodes[parent]["Rmin"]
        left = nodes[index]["child"]
        if left != -1:
            stack += [(left + 1, index, 0), (left, index, 1)]
    return bounds


def reference_validity(nodes, bounds, bb_indices, bb_coords):  # validate_node_bounds, query.py:624-664
    def contained(a, b):
        return a[0] >= b[0] and a[1] <= b[1] and a[2] >= b[2] and a[3] <= b[3]

    valid = np.zeros(len(nodes), dtype=bool)
    stack = [0]
    while stack:
        index = stack.pop()
        node = nodes[index]
        if node["child"] == -1:
            cells = bb_indices[node["ptr"] : node["ptr"] + node["size"]]
            valid[index] = all(contained(bb_coords[c], bounds[index]) for c in cells)
            continue
        left = node["child"]
        valid[index] = contained(bounds[left], bounds[index]) and contained(bounds[left + 1], bounds[index])
        stack += [left + 1, left]
    return valid


@pytest.mark.parametrize("mesh", ["delaunay", "quads", "single"])
def test_diagnostics_match_the_reference_loops(mesh):
    if mesh == "delaunay":
        vertices, faces = delaunay_mesh(1_500, seed=4)
    elif mesh == "quads":
        vertices, faces = quad_mesh(23, 17)
    else:
        vertices, faces = np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0]]), np.array([[0, 1, 2]])
    ref = oracle.CellTree2d(vertices, faces, -1, cells_per_leaf=3)
    tree = MirrorsOnly(ref)
    rng = np.random.default_rng(1)
    for mutation in range(4):
        bounds = tree.node_bounds
        want = reference_bounds(tree.nodes, tree.bbox)
        assert np.array_equal(bounds, want)
        assert np.array_equal(tree.validate_node_bounds(), reference_validity(tree.nodes, want, tree.bb_indices, tree.bb_coords))
        d = tree.to_dict_of_lists()
        assert list(d) == list(range(len(tree.nodes)))
        assert all(v == ([] if c == -1 else [c, c + 1]) for v, c in zip(d.values(), tree.nodes["child"]))
        inner = np.flatnonzero(tree.nodes["child"] != -1)
        if len(inner) == 0:
            break
        k = rng.choice(inner)  # squeeze a plane: some subtree is no longer contained
        tree.nodes[k]["Lmax"] -= 0.05 * (mutation + 1)
        tree.nodes[k]["Rmin"] += 0.03 * (mutation + 1)
    if mesh != "single":
        assert not tree.validate_node_bounds().all()
