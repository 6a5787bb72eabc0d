// traverse.cuh -- per-query stack traversal of the cell tree.
//
// The reference keeps an explicit stack of node indices (query.py:63-107, 177-244, 357-455), pushes
// child A then child B and pops B first.  Here a Cursor sits on the node to visit next and only the
// deferred sibling goes to the per-thread stack: "push A, push B" becomes "stack <- A, cursor -> B".
// The visiting order -- and with it the first-hit result of locate_points and the emission order of
// box / edge pairs -- is exactly the reference's.  The nodes are read through the treelets of common.cuh
// (three levels per 128-byte line); a stack entry is a node handle (treelet << 3 | slot).
//
// The stack is a per-thread local array: it is thread-interleaved in local memory, so the 32 lanes of a
// warp touch one 128-byte line per slot, and it stays in L1.  Live depth is at most (tree depth - 1):
// one deferred sibling per level of the current path.  ct_tree.depth is checked against STACK_CAP before
// any launch (CT_ERR_DEPTH instead of silent truncation).
//
// Points walk in two nested loops (descend; then the leaf), their lanes move in step.  Boxes and segments walk
// in one loop whose step has as few divergent paths as possible (the outcome of an inner node becomes selects and
// a predicated push): their lanes are at different kinds of node most of the time.
#pragma once

#include "geometry.cuh"

namespace ct {

constexpr int STACK_CAP = 64;

// ---- the descent, shared by the four traversals ---------------------------------------------------------------
// Cursor on one binary node = (treelet, slot), see common.cuh.  The node's 16-byte slot is loaded when the cursor
// moves; the treelet header only when it enters another treelet, i.e. every third level or after a pop -- header and
// root slot then come from the same 32-byte sector.
constexpr uint32_t ROOT_HANDLE = 1u;  // treelet 0, slot 1

struct Cursor {
    uint32_t handle;
    uint32_t child_base, meta, child_off;
    double2 plane;  // inner: (Lmax, Rmin); leaf: bits of {ptr, size, id0, id1}
};
CT_DEV void cursor_load_plane(Cursor &c, const char *__restrict__ base) {
    c.plane = __ldg(reinterpret_cast<const double2 *>(base + ((size_t)c.handle << 4)));
}
CT_DEV void cursor_enter(Cursor &c, const char *__restrict__ base, uint32_t handle) {
    c.handle = handle;
    const char *slot = base + ((size_t)handle << 4);
    // the header is the first slot of the same (128-byte aligned) line
    const uint4 header = __ldg(reinterpret_cast<const uint4 *>(reinterpret_cast<uintptr_t>(slot) & ~(uintptr_t)127));
    c.plane = __ldg(reinterpret_cast<const double2 *>(slot));
    c.child_base = header.x;
    c.meta = header.y;
    c.child_off = header.z;
}
CT_DEV bool cursor_is_leaf(const Cursor &c) { return (c.meta >> (8u + (c.handle & 7u))) & 1u; }
CT_DEV bool cursor_dim(const Cursor &c) { return (c.meta >> (c.handle & 7u)) & 1u; }
// handles of the two children of the inner node under the cursor
CT_DEV void cursor_children(const Cursor &c, uint32_t &left, uint32_t &right) {
    const uint32_t q = c.handle & 7u;
    if (q < 4u) {
        left = c.handle + q;  // same treelet, slot 2q
        right = left + 1u;
    } else {
        left = ((c.child_base + ((c.child_off >> (4u * (q - 4u))) & 15u)) << 3) | 1u;
        right = left + 8u;
    }
}
// move to a child of the node under the cursor
CT_DEV void cursor_descend(Cursor &c, const char *__restrict__ base, uint32_t child) {
    if ((c.handle & 7u) < 4u) {
        c.handle = child;
        cursor_load_plane(c, base);
    } else {
        cursor_enter(c, base, child);
    }
}
// where the descent of a point query starts (common.cuh: EntryGrid)
CT_DEV uint32_t entry_handle(const EntryGrid &g, P2 p) {
    if (g.handle == nullptr) return ROOT_HANDLE;
    const int shift = 16 - g.bits;
    const uint32_t cx = grid_coord(p.x, g.xmin, g.sx) >> shift, cy = grid_coord(p.y, g.ymin, g.sy) >> shift;
    const bool inside = p.x > __ldg(g.lo + cx) && p.y > __ldg(g.lo + (1 << g.bits) + cy);  // false for NaN
    return inside ? __ldg(g.handle + ((cy << g.bits) | cx)) : ROOT_HANDLE;
}

CT_DEV int4 cursor_leaf(const Cursor &c) {  // {ptr, size, id0, id1}
    const long long a = __double_as_longlong(c.plane.x), b = __double_as_longlong(c.plane.y);
    return make_int4((int)(a & 0xffffffffLL), (int)(a >> 32), (int)(b & 0xffffffffLL), (int)(b >> 32));
}
// k-th element of a leaf: the first two ids travel with the node, the rest come from bb_indices
CT_DEV int leaf_element(const int4 &leaf, const int32_t *__restrict__ bb_indices, int k) {
    return k == 0 ? leaf.z : (k == 1 ? leaf.w : __ldg(bb_indices + leaf.x + k));
}

// ---- locate_point, query.py:63-107 ----------------------------------------------------------------------
// Returns the element index of the first face (in DFS order) that contains the point, -1 if none.
// Two nested loops: the inner one only descends (a handful of registers), the outer one tests the cells of the
// leaf it arrived at (query.py:72-85) -- so that the register allocation of the descent is not weighed down by
// the point-in-polygon test.
template <int MAXV>
CT_DEV int locate_point(const TreeView &t, P2 p, double tolerance) {
    uint32_t stack[STACK_CAP];
    int sp = 0;
    const char *base = reinterpret_cast<const char *>(t.treelets);
    Cursor c;
    cursor_enter(c, base, entry_handle(t.entry, p));
    while (true) {
        while (!cursor_is_leaf(c)) {
            const double Lmax = c.plane.x, Rmin = c.plane.y;
            const double pd = cursor_dim(c) ? p.y : p.x;
            const bool left = pd <= Lmax;
            const bool right = pd >= Rmin;
            uint32_t left_handle, right_handle;
            cursor_children(c, left_handle, right_handle);
            if (left && right) {
                // nearer-plane heuristic, query.py:93-101: the child pushed LAST is visited first
                const bool right_first = (Lmax - pd) < (pd - Rmin);
                stack[sp++] = right_first ? left_handle : right_handle;
                cursor_descend(c, base, right_first ? right_handle : left_handle);
            } else if (left) {
                cursor_descend(c, base, left_handle);
            } else if (right) {
                cursor_descend(c, base, right_handle);
            } else {
                if (sp == 0) return -1;
                cursor_enter(c, base, stack[--sp]);
            }
        }
        const int4 leaf = cursor_leaf(c);
        for (int k = 0; k < leaf.y; k++) {
            int bbox_index = leaf_element(leaf, t.bb_indices, k);
            Poly<MAXV> poly;
            load_polygon<MAXV>(t.elements, t.M, bbox_index, t.elem_xy, poly);
            if (point_in_polygon_or_on_edge(p, poly, tolerance)) return bbox_index;
        }
        if (sp == 0) return -1;
        cursor_enter(c, base, stack[--sp]);
    }
}

// ---- locate_point_on_edge, query.py:121-165 ---------------------------------------------------------------
CT_DEV int locate_point_on_edge(const TreeView &t, P2 p, double tolerance) {
    uint32_t stack[STACK_CAP];
    int sp = 0;
    const char *base = reinterpret_cast<const char *>(t.treelets);
    Cursor c;
    cursor_enter(c, base, entry_handle(t.entry, p));
    while (true) {
        while (!cursor_is_leaf(c)) {
            const double Lmax = c.plane.x, Rmin = c.plane.y;
            const double pd = cursor_dim(c) ? p.y : p.x;
            const bool left = pd <= Lmax;
            const bool right = pd >= Rmin;
            uint32_t left_handle, right_handle;
            cursor_children(c, left_handle, right_handle);
            if (left && right) {
                const bool right_first = (Lmax - pd) < (pd - Rmin);
                stack[sp++] = right_first ? left_handle : right_handle;
                cursor_descend(c, base, right_first ? right_handle : left_handle);
            } else if (left) {
                cursor_descend(c, base, left_handle);
            } else if (right) {
                cursor_descend(c, base, right_handle);
            } else {
                if (sp == 0) return -1;
                cursor_enter(c, base, stack[--sp]);
            }
        }
        const int4 leaf = cursor_leaf(c);
        for (int k = 0; k < leaf.y; k++) {
            int bbox_index = leaf_element(leaf, t.bb_indices, k);
            double2 v0 = __ldg(t.elem_xy + 2 * (int64_t)bbox_index);
            double2 v1 = __ldg(t.elem_xy + 2 * (int64_t)bbox_index + 1);
            if (point_on_edge(p, P2{v0.x, v0.y}, P2{v1.x, v1.y}, tolerance)) return bbox_index;
        }
        if (sp == 0) return -1;
        cursor_enter(c, base, stack[--sp]);
    }
}

// ---- locate_box, query.py:177-244 -------------------------------------------------------------------------
// Emit(bbox_index) is called for every leaf cell whose bounding box strictly overlaps `box`, in the
// reference's order (right subtree first: it pushes left then right and pops right).  Returns the count.
// A box visits ~90 nodes and the lanes of a warp are at different kinds of node most of the time (ncu: 9 of 32
// lanes active per instruction), so the step is written with as few divergent paths as possible: one for a leaf,
// one for an inner node whose outcome (left / right / both / neither) is turned into selects and a predicated push,
// and a common tail that pops if needed and always loads header + slot of the next node.
template <typename Emit>
CT_DEV int locate_box(const TreeView &t, const Box4 &box, Emit emit) {
    Box4 tree_bbox{t.bbox[0], t.bbox[1], t.bbox[2], t.bbox[3]};
    if (!boxes_intersect(box, tree_bbox)) return 0;
    uint32_t stack[STACK_CAP];
    int sp = 0;
    int count = 0;
    const char *base = reinterpret_cast<const char *>(t.treelets);
    Cursor c;
    cursor_enter(c, base, ROOT_HANDLE);
    while (true) {
        uint32_t next = 0;
        bool pop;
        if (cursor_is_leaf(c)) {
            const int4 leaf = cursor_leaf(c);
            for (int k = 0; k < leaf.y; k++) {
                int bbox_index = leaf_element(leaf, t.bb_indices, k);
                Box4 leaf_box = load_box(t.bb_coords, bbox_index);
                if (boxes_intersect(box, leaf_box)) {
                    emit(count, bbox_index);
                    count++;
                }
            }
            pop = true;
        } else {
            const bool dim = cursor_dim(c);
            const double bmin = dim ? box.ymin : box.xmin;
            const double bmax = dim ? box.ymax : box.xmax;
            const bool left = bmin <= c.plane.x;
            const bool right = bmax >= c.plane.y;
            uint32_t left_handle, right_handle;
            cursor_children(c, left_handle, right_handle);
            if (left && right) stack[sp++] = left_handle;
            next = right ? right_handle : left_handle;
            pop = !(left || right);
        }
        if (pop) {
            if (sp == 0) return count;
            next = stack[--sp];
        }
        cursor_enter(c, base, next);
    }
}

// ---- per-candidate tests of locate_edge -------------------------------------------------------------------
// compute_edge_face_intersect, query.py:309-328, in its two halves: the Cohen-Sutherland clip against the cell's
// bounding box decides whether the cell is looked at at all (its clipped points are not used further) ...
CT_DEV bool edge_face_prefilter(const TreeView &t, int bbox_index, P2 a, P2 b) {
    Box4 box = load_box(t.bb_coords, bbox_index);
    P2 c, d;
    return cohen_sutherland_line_box_clip(a, b, box, c, d) != 0;
}
// ... and the Cyrus-Beck clip against the cell's polygon gives the intersection
template <int MAXV>
CT_DEV bool edge_face_clip(const TreeView &t, int bbox_index, P2 a, P2 b, P2 &c, P2 &d) {
    Box4 box = load_box(t.bb_coords, bbox_index);
    Poly<MAXV> polygon;
    load_polygon<MAXV>(t.elements, t.M, bbox_index, t.elem_xy, polygon);
    double tolerance = nb_max(MIN_TOLERANCE, TOLERANCE_FACTOR * nb_max(box.xmax - box.xmin, box.ymax - box.ymin));
    return cyrus_beck_line_polygon_clip<MAXV>(a, b, polygon, tolerance, c, d);
}
template <int MAXV>
CT_DEV bool edge_face_intersect(const TreeView &t, int bbox_index, P2 a, P2 b, P2 &c, P2 &d) {
    if (!edge_face_prefilter(t, bbox_index, a, b)) return false;
    return edge_face_clip<MAXV>(t, bbox_index, a, b, c, d);
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
// Which children of the inner node under the cursor the segment a -> b (V = b - a) may reach: the parametric test of
// the planes Lmax / Rmin along the segment, query.py:407-440.
//
// The reference first sets left = (dx_left >= 0), right = (dx_right <= 0) and then, where such a flag is set and dx != 0,
// replaces it by a test of the parameter t = dx_left / dx (forward, dx > 0: t >= 0; backward, dx < 0: 1 - t >= 0; for the
// right plane: t <= 1, resp. 1 - t <= 1).  With FINITE operands those second tests cannot fail:
//   forward:  dx_left >= 0, dx > 0  =>  t is +-0 or positive (or +0 by underflow)     =>  t >= 0;
//             dx_right <= 0          =>  t is +-0 or negative (-inf by overflow)        =>  t <= 1;
//   backward: dx_left >= 0, dx < 0  =>  t <= 0 or -0  =>  1 - t >= 1                   =>  1 - t >= 0;
//             dx_right <= 0          =>  t >= 0 or +0 (+inf by overflow)  =>  1 - t <= 1.
// (IEEE division gives the quotient the XOR of the signs, and -0.0 >= 0.0 holds.)  Only an infinite operand can change
// the outcome (inf / inf = NaN fails every comparison), so the two divisions -- 8 % of the instructions of
// intersect_edges' first pass (profiles/r01_edges_cooperative_ncu.txt) -- are formed on that path alone.
CT_DEV void edge_plane_test(const Cursor &cur, P2 a, P2 b, P2 V, bool &left, bool &right) {
    const bool dim = cursor_dim(cur);
    const double Lmax = cur.plane.x, Rmin = cur.plane.y;
    const double dx = dim ? V.y : V.x;
    const double a_d = dim ? a.y : a.x;
    const double b_d = dim ? b.y : b.x;
    const bool forward = dx > 0.0, backward = dx < 0.0;
    const double dx_left = Lmax - (forward ? a_d : b_d);
    const double dx_right = Rmin - (forward ? b_d : a_d);
    left = dx_left >= 0.0;
    right = dx_right <= 0.0;
    const bool finite = fabs(dx) <= FLOAT_MAX && fabs(dx_left) <= FLOAT_MAX && fabs(dx_right) <= FLOAT_MAX;  // false for NaN
    if (!finite) {
        const double t_left = dx_left / dx, t_right = dx_right / dx;
        const bool left_t = forward ? (t_left >= 0.0) : (backward ? ((1.0 - t_left) >= 0.0) : true);
        const bool right_t = forward ? (t_right <= 1.0) : (backward ? ((1.0 - t_right) <= 1.0) : true);
        left = left && left_t;
        right = right && right_t;
    }
}

// MAXV == 0 selects the edge-edge test (EdgeCellTree2d), otherwise the edge-face test with that polygon bound.
template <int MAXV>
CT_DEV bool edge_cell_intersect(const TreeView &t, int bbox_index, P2 a, P2 b, P2 &c, P2 &d) {
    if constexpr (MAXV == 0) return edge_edge_intersect(t, bbox_index, a, b, c, d);
    else return edge_face_intersect<MAXV>(t, bbox_index, a, b, c, d);
}

// One thread walks one segment and tests every candidate as it meets it (the second traversal of the count -> fill
// scheme; the first pass is the warp-cooperative kernel of edges.cu).
template <int MAXV, typename Emit>
CT_DEV int locate_edge(const TreeView &t, P2 a, P2 b, Emit emit) {
    {
        P2 c, d;
        Box4 tree_bbox{t.bbox[0], t.bbox[1], t.bbox[2], t.bbox[3]};
        if (!cohen_sutherland_line_box_clip(a, b, tree_bbox, c, d)) return 0;
    }
    P2 V = to_vector(a, b);
    uint32_t stack[STACK_CAP];
    int sp = 0;
    int count = 0;
    const char *base = reinterpret_cast<const char *>(t.treelets);
    Cursor cur;
    cursor_enter(cur, base, ROOT_HANDLE);
    while (true) {
        uint32_t next = 0;
        bool pop;
        if (cursor_is_leaf(cur)) {
            const int4 leaf = cursor_leaf(cur);
            for (int k = 0; k < leaf.y; k++) {
                int bbox_index = leaf_element(leaf, t.bb_indices, k);
                P2 c, d;
                if (edge_cell_intersect<MAXV>(t, bbox_index, a, b, c, d)) {
                    emit(count, bbox_index, c, d);
                    count++;
                }
            }
            pop = true;
        } else {
            bool left, right;
            edge_plane_test(cur, a, b, V, left, right);
            uint32_t left_handle, right_handle;
            cursor_children(cur, left_handle, right_handle);
            if (left && right) stack[sp++] = left_handle;
            next = right ? right_handle : left_handle;
            pop = !(left || right);
        }
        if (pop) {
            if (sp == 0) return count;
            next = stack[--sp];
        }
        cursor_enter(cur, base, next);
    }
}

}  // namespace ct
