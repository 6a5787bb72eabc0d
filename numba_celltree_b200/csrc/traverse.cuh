// traverse.cuh -- per-query stack traversal of the cell tree (one thread per query).
//
// The reference keeps an explicit stack of node indices (query.py:63-107, 177-244, 357-455), pushes
// child A then child B and pops B first.  Here the node to visit next lives in a register and only the
// deferred sibling goes to the per-thread stack: "push A, push B" becomes "stack <- A, next = B".
// The visiting order -- and with it the first-hit result of locate_points and the emission order of
// box / edge pairs -- is exactly the reference's.
//
// The stack is a per-thread local array: it is thread-interleaved in local memory, so the 32 lanes of a
// warp touch one 128-byte line per slot, and it stays in L1.  Live depth is at most (tree depth - 1):
// one deferred sibling per level of the current path.  ct_tree.depth is checked against STACK_CAP before
// any launch (CT_ERR_DEPTH instead of silent truncation).
#pragma once

#include "geometry.cuh"

namespace ct {

constexpr int STACK_CAP = 64;

constexpr int LEAF_INLINE = 4;  // element ids stored in a leaf's (unused) plane fields

// k-th element of a leaf: the first LEAF_INLINE ids travel with the node, the rest come from bb_indices
CT_DEV int leaf_element(const Node32 &node, const int32_t *__restrict__ bb_indices, int k) {
    if (k < LEAF_INLINE) {
        long long bits = __double_as_longlong(k < 2 ? node.Lmax : node.Rmin);
        return (k & 1) ? (int)(bits >> 32) : (int)(bits & 0xffffffffLL);
    }
    return __ldg(bb_indices + node.ptr + k);
}

CT_DEV Node32 load_node(const Node32 *__restrict__ nodes, int idx) {
    // one 32-byte sector, two 16-byte read-only loads
    const double2 *p = reinterpret_cast<const double2 *>(nodes + idx);
    double2 lr = __ldg(p);
    int4 m = __ldg(reinterpret_cast<const int4 *>(p + 1));
    Node32 n;
    n.Lmax = lr.x;
    n.Rmin = lr.y;
    n.child = m.x;
    n.ptr = m.y;
    n.size = m.z;
    n.dim = m.w;
    return n;
}

// ---- locate_point, query.py:63-107 ----------------------------------------------------------------------
// Returns the element index of the first face (in DFS order) that contains the point, -1 if none.
// On a hit `poly` holds that face's vertices (used by the fused barycentric weights).
template <int MAXV>
CT_DEV int locate_point(const TreeView &t, P2 p, double tolerance, Poly<MAXV> &poly) {
    int stack[STACK_CAP];
    int sp = 0;
    int node_index = 0;
    while (true) {
        Node32 node = load_node(t.nodes, node_index);
        bool pop = false;
        if (node.child == -1) {
            for (int k = 0; k < node.size; k++) {
                int bbox_index = leaf_element(node, t.bb_indices, k);
                load_polygon<MAXV>(t.elements, t.M, bbox_index, t.elem_xy, poly);
                if (point_in_polygon_or_on_edge(p, poly, tolerance)) return bbox_index;
            }
            pop = true;
        } else {
            double pd = node.dim ? p.y : p.x;
            bool left = pd <= node.Lmax;
            bool right = pd >= node.Rmin;
            int left_child = node.child;
            int right_child = left_child + 1;
            if (left && right) {
                // nearer-plane heuristic, query.py:93-101: the child pushed LAST is visited first
                if ((node.Lmax - pd) < (pd - node.Rmin)) {
                    stack[sp++] = left_child;
                    node_index = right_child;
                } else {
                    stack[sp++] = right_child;
                    node_index = left_child;
                }
            } else if (left) {
                node_index = left_child;
            } else if (right) {
                node_index = right_child;
            } else {
                pop = true;
            }
        }
        if (pop) {
            if (sp == 0) return -1;
            node_index = stack[--sp];
        }
    }
}

// ---- locate_point_on_edge, query.py:121-165 ---------------------------------------------------------------
CT_DEV int locate_point_on_edge(const TreeView &t, P2 p, double tolerance) {
    int stack[STACK_CAP];
    int sp = 0;
    int node_index = 0;
    while (true) {
        Node32 node = load_node(t.nodes, node_index);
        bool pop = false;
        if (node.child == -1) {
            for (int k = 0; k < node.size; k++) {
                int bbox_index = leaf_element(node, t.bb_indices, k);
                double2 v0 = __ldg(t.elem_xy + 2 * (int64_t)bbox_index);
                double2 v1 = __ldg(t.elem_xy + 2 * (int64_t)bbox_index + 1);
                if (point_on_edge(p, P2{v0.x, v0.y}, P2{v1.x, v1.y}, tolerance)) return bbox_index;
            }
            pop = true;
        } else {
            double pd = node.dim ? p.y : p.x;
            bool left = pd <= node.Lmax;
            bool right = pd >= node.Rmin;
            int left_child = node.child;
            int right_child = left_child + 1;
            if (left && right) {
                if ((node.Lmax - pd) < (pd - node.Rmin)) {
                    stack[sp++] = left_child;
                    node_index = right_child;
                } else {
                    stack[sp++] = right_child;
                    node_index = left_child;
                }
            } else if (left) {
                node_index = left_child;
            } else if (right) {
                node_index = right_child;
            } else {
                pop = true;
            }
        }
        if (pop) {
            if (sp == 0) return -1;
            node_index = stack[--sp];
        }
    }
}

// ---- locate_box, query.py:177-244 -------------------------------------------------------------------------
// Emit(bbox_index) is called for every leaf cell whose bounding box strictly overlaps `box`, in the
// reference's order (right subtree first: it pushes left then right and pops right).  Returns the count.
template <typename Emit>
CT_DEV int locate_box(const TreeView &t, const Box4 &box, Emit emit) {
    Box4 tree_bbox{t.bbox[0], t.bbox[1], t.bbox[2], t.bbox[3]};
    if (!boxes_intersect(box, tree_bbox)) return 0;
    int stack[STACK_CAP];
    int sp = 0;
    int node_index = 0;
    int count = 0;
    while (true) {
        Node32 node = load_node(t.nodes, node_index);
        bool pop = false;
        if (node.child == -1) {
            for (int k = 0; k < node.size; k++) {
                int bbox_index = leaf_element(node, t.bb_indices, k);
                Box4 leaf_box = load_box(t.bb_coords, bbox_index);
                if (boxes_intersect(box, leaf_box)) {
                    emit(count, bbox_index);
                    count++;
                }
            }
            pop = true;
        } else {
            double bmin = node.dim ? box.ymin : box.xmin;
            double bmax = node.dim ? box.ymax : box.xmax;
            bool left = bmin <= node.Lmax;
            bool right = bmax >= node.Rmin;
            int left_child = node.child;
            int right_child = left_child + 1;
            if (left && right) {
                stack[sp++] = left_child;
                node_index = right_child;
            } else if (left) {
                node_index = left_child;
            } else if (right) {
                node_index = right_child;
            } else {
                pop = true;
            }
        }
        if (pop) {
            if (sp == 0) return count;
            node_index = stack[--sp];
        }
    }
}

// ---- per-candidate tests of locate_edge -------------------------------------------------------------------
// compute_edge_face_intersect, query.py:309-328
template <int MAXV>
CT_DEV bool edge_face_intersect(const TreeView &t, int bbox_index, P2 a, P2 b, P2 &c, P2 &d) {
    Box4 box = load_box(t.bb_coords, bbox_index);
    bool intersects = cohen_sutherland_line_box_clip(a, b, box, c, d) != 0;
    if (intersects) {
        Poly<MAXV> polygon;
        load_polygon<MAXV>(t.elements, t.M, bbox_index, t.elem_xy, polygon);
        double tolerance = nb_max(MIN_TOLERANCE, TOLERANCE_FACTOR * nb_max(box.xmax - box.xmin, box.ymax - box.ymin));
        intersects = cyrus_beck_line_polygon_clip<MAXV>(a, b, polygon, tolerance, c, d);
    }
    return intersects;
}

// compute_edge_edge_intersect, query.py:292-306
CT_DEV bool edge_edge_intersect(const TreeView &t, int bbox_index, P2 a, P2 b, P2 &c, P2 &d) {
    double2 p = __ldg(t.elem_xy + 2 * (int64_t)bbox_index);
    double2 q = __ldg(t.elem_xy + 2 * (int64_t)bbox_index + 1);
    bool intersects = lines_intersect(a, b, P2{p.x, p.y}, P2{q.x, q.y}, c);
    d = c;
    return intersects;
}

// ---- locate_edge, query.py:357-455 --------------------------------------------------------------------------
// MAXV == 0 selects the edge-edge test (EdgeCellTree2d), otherwise the edge-face test with that polygon bound.
template <int MAXV, typename Emit>
CT_DEV int locate_edge(const TreeView &t, P2 a, P2 b, Emit emit) {
    {
        P2 c, d;
        Box4 tree_bbox{t.bbox[0], t.bbox[1], t.bbox[2], t.bbox[3]};
        if (!cohen_sutherland_line_box_clip(a, b, tree_bbox, c, d)) return 0;
    }
    P2 V = to_vector(a, b);
    int stack[STACK_CAP];
    int sp = 0;
    int node_index = 0;
    int count = 0;
    while (true) {
        Node32 node = load_node(t.nodes, node_index);
        bool pop = false;
        if (node.child == -1) {
            for (int k = 0; k < node.size; k++) {
                int bbox_index = leaf_element(node, t.bb_indices, k);
                P2 c, d;
                bool intersects;
                if constexpr (MAXV == 0) intersects = edge_edge_intersect(t, bbox_index, a, b, c, d);
                else intersects = edge_face_intersect<MAXV>(t, bbox_index, a, b, c, d);
                if (intersects) {
                    emit(count, bbox_index, c, d);
                    count++;
                }
            }
            pop = true;
        } else {
            // parametric test of the planes Lmax / Rmin along the segment, query.py:407-440
            double dx = node.dim ? V.y : V.x;
            double a_d = node.dim ? a.y : a.x;
            double b_d = node.dim ? b.y : b.x;
            double dx_left, dx_right;
            if (dx > 0.0) {
                dx_left = node.Lmax - a_d;
                dx_right = node.Rmin - b_d;
            } else {
                dx_left = node.Lmax - b_d;
                dx_right = node.Rmin - a_d;
            }
            bool left = dx_left >= 0.0;
            bool right = dx_right <= 0.0;
            if (dx > 0.0) {
                if (left) left = (dx_left / dx) >= 0.0;
                if (right) right = (dx_right / dx) <= 1.0;
            } else if (dx < 0.0) {
                if (left) left = (1.0 - (dx_left / dx)) >= 0.0;
                if (right) right = (1.0 - (dx_right / dx)) <= 1.0;
            }
            int left_child = node.child;
            int right_child = left_child + 1;
            if (left && right) {
                stack[sp++] = left_child;
                node_index = right_child;
            } else if (left) {
                node_index = left_child;
            } else if (right) {
                node_index = right_child;
            } else {
                pop = true;
            }
        }
        if (pop) {
            if (sp == 0) return count;
            node_index = stack[--sp];
        }
    }
}

}  // namespace ct
