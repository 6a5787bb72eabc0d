"""
Host-visible types and constants, byte-compatible with the reference's (constants.py:27-147): users and
tests address ``tree.nodes`` fields by name and type-check ``tree.celltree_data``.
"""

from typing import NamedTuple

import numpy as np

IntDType = np.intp
FloatDType = np.float64
IntArray = np.ndarray
FloatArray = np.ndarray
BoolArray = np.ndarray
NodeArray = np.ndarray


class CellTreeData(NamedTuple):
    elements: IntArray
    vertices: FloatArray
    nodes: NodeArray
    bb_indices: IntArray
    bb_coords: FloatArray
    bbox: FloatArray
    cells_per_leaf: int


# 41 bytes, packed: the image the C-ABI's ct_node41 reads and writes.
NodeDType = np.dtype(
    [
        ("child", IntDType),  # index of left child; right child is child + 1; -1 for a leaf
        ("Lmax", FloatDType),
        ("Rmin", FloatDType),
        ("ptr", IntDType),  # into bb_indices
        ("size", IntDType),
        ("dim", bool),  # False = x, True = y
    ]
)
assert NodeDType.itemsize == 41

NDIM = 2
MAX_N_VERTEX = 32
FILL_VALUE = -1
MIN_TOLERANCE = 1e-15
TOLERANCE_FACTOR = 1e-12
