// traverse.cuh -- per-query stack traversal of the cell tree.
//
// The reference keeps an explicit stack of node indices (query.py:63-107, 177-244, 357-455), pushes
// child A then child B and pops B first.  Here a Cursor sits on the node to visit next and only the
// deferred sibling goes to the per-thread stack: "push A, push B" becomes "stack <- A, cursor -> B".
// The visiting order -- and with it the first-hit result of locate_points and the emission order of
// box / edge pairs -- is exactly the reference's.  The nodes are read through the treelets of common.cuh
// (three levels per 128-byte line); a stack entry is a node handle (treelet << 3 | slot).
//
// The stack's first STACK_CAP entries are a per-thread local array: it is thread-interleaved in local memory, so the
// 32 lanes of a warp touch one 128-byte line per slot, and it stays in L1.  Live depth is at most (tree depth - 1):
// one deferred sibling per level of the current path; deeper trees continue in an overflow slab (Stack, below).
//
// Points walk in two nested loops (descend; then the leaf), their lanes move in step.  Boxes and segments walk
// in one loop whose step has as few divergent paths as possible (the outcome of an inner node becomes selects and
// a predicated push): their lanes are at different kinds of node most of the time.
#pragma once

#include "geometry.cuh"

namespace ct {

constexpr uint32_t ROOT_HANDLE_VALUE = 1u;
constexpr int STACK_CAP = 64;

// The reference's stack grows when it is full (utils.py:35-44, constants.py:136), so a tree of any depth is walked.
// Here the first STACK_CAP entries are a per-thread array -- enough for every tree of at most STACK_CAP levels, i.e. for
// every sane mesh -- and a thread whose stack outgrows them continues in a column of the tree's overflow slab
// (common.cuh: DeepStacks; a cold, out-of-line path).
// entry k of the overflow column of a thread (taken from the slab on first use); returns the column.  Out of line and by
// value: the cold path must not force the caller's stack object into memory.
static __device__ __noinline__ uint32_t *deep_stack_store(DeepStacks d, uint32_t *column, int k, uint32_t v) {
    if (d.slab == nullptr) return column;  // unreachable for a tree whose depth the host has checked
    if (column == nullptr) {
        int c = atomicAdd(d.state, 1);
        if (c >= d.slots) {  // no column left: the host reports CT_ERR_DEPTH, nothing of this launch is used
            d.state[1] = 1;
            c = 0;
        }
        column = d.slab + c;
    }
    column[(size_t)k * d.slots] = v;
    return column;
}

// DEEP = false: the tree has at most STACK_CAP levels (the host checks), the stack is the plain per-thread array.
// DEEP = true: kernels launched for deeper trees only; their stacks continue in the overflow slab.
// The per-thread array is declared by the caller (CT_STACK) and only pointed to from here: with the array inside this
// struct the compiler keeps the WHOLE struct in local memory (the array is indexed dynamically), stack pointer included,
// and every push / pop / empty test becomes a local load (measured: +6 % on the box and segment walks).
template <bool DEEP>
struct StackT {
    uint32_t *local;
    uint32_t *deep = nullptr;
    int sp = 0;
    CT_DEV explicit StackT(uint32_t *memory) : local(memory) {}
    CT_DEV bool empty() const { return sp == 0; }
    CT_DEV void push(const TreeView &t, uint32_t v) {
        if constexpr (DEEP) {
            if (sp < STACK_CAP) local[sp] = v;
            else deep = deep_stack_store(t.deep, deep, sp - STACK_CAP, v);
            sp++;
        } else {
            local[sp++] = v;
        }
    }
    CT_DEV uint32_t pop(const TreeView &t) {
        --sp;
        if constexpr (DEEP) {
            if (sp >= STACK_CAP) return deep != nullptr ? deep[(size_t)(sp - STACK_CAP) * t.deep.slots] : ROOT_HANDLE_VALUE;
        }
        return local[sp];
    }
};

#define CT_STACK(name, DEEP_FLAG)          \
    uint32_t name##_memory[STACK_CAP];     \
    StackT<DEEP_FLAG> name(name##_memory)

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
    // the three reads are independent (grid_coord() stays inside the table for any input), so they travel together
    const double lo_x = __ldg(g.lo + cx), lo_y = __ldg(g.lo + (1 << g.bits) + cy);
    const uint32_t handle = __ldg(g.handle + ((cy << g.bits) | cx));
    const bool inside = p.x > lo_x && p.y > lo_y;  // false for NaN
    return inside ? handle : ROOT_HANDLE;
}

CT_DEV int4 cursor_leaf(const Cursor &c) {  // {ptr, size, id0, id1}
    const long long a = __double_as_longlong(c.plane.x), b = __double_as_longlong(c.plane.y);
    return make_int4((int)(a & 0xffffffffLL), (int)(a >> 32), (int)(b & 0xffffffffLL), (int)(b >> 32));
}
// Every element of a leaf in its order: the first two ids travel with the node (no run-time selection among them: a
// `k == 0 ? z : k == 1 ? w : load` chain inside a loop was 16 % of the box walk's instructions, at five active lanes), the
// rest come from bb_indices.  `visit` returns true to stop early.
template <typename Visit>
CT_DEV bool for_each_leaf_element(const int4 &leaf, const int32_t *__restrict__ bb_indices, Visit visit) {
    if (leaf.y > 0 && visit(leaf.z)) return true;
    if (leaf.y > 1 && visit(leaf.w)) return true;
    for (int k = 2; k < leaf.y; k++)
        if (visit(__ldg(bb_indices + leaf.x + k))) return true;
    return false;
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
// `Probe` counts what a walk touches (NoProbe: nothing, compiled away); ct_locate_points_stats uses it to state the bytes
// the traversal as built has to move per query.
struct NoProbe {
    CT_DEV void slot() {}
    CT_DEV void header() {}
    CT_DEV void cell() {}
    CT_DEV void push() {}
    CT_DEV void entry(bool) {}
};
struct CountingProbe {
    unsigned slots = 0, headers = 0, cells = 0, pushes = 0, entries = 0;
    CT_DEV void slot() { slots++; }
    CT_DEV void header() { headers++; }
    CT_DEV void cell() { cells++; }
    CT_DEV void push() { pushes++; }
    CT_DEV void entry(bool from_grid) { entries += from_grid ? 1u : 0u; }
};

template <int MAXV, typename Probe = NoProbe, bool DEEP = false, bool FILTER = true>
CT_DEV int locate_point(const TreeView &t, P2 p, double tolerance, Probe *probe = nullptr) {
    CT_STACK(stack, DEEP);
    const char *base = reinterpret_cast<const char *>(t.treelets);
    Cursor c;
    Probe none;
    Probe &pr = probe ? *probe : none;
    const double margin = outside_margin(tolerance);
    const bool x_sides = outside_x_sides_allowed(p);
    // a move within a treelet reads the node's slot; entering a treelet reads its header as well
    auto descend = [&](uint32_t child) {
        pr.slot();
        if ((c.handle & 7u) >= 4u) pr.header();
        cursor_descend(c, base, child);
    };
    auto enter = [&](uint32_t handle) {
        pr.slot();
        pr.header();
        cursor_enter(c, base, handle);
    };
    {
        const uint32_t first = entry_handle(t.entry, p);
        pr.entry(first != ROOT_HANDLE);
        enter(first);
    }
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
                pr.push();
                stack.push(t, right_first ? left_handle : right_handle);
                descend(right_first ? right_handle : left_handle);
            } else if (left) {
                descend(left_handle);
            } else if (right) {
                descend(right_handle);
            } else {
                if (stack.empty()) return -1;
                enter(stack.pop(t));
            }
        }
        const int4 leaf = cursor_leaf(c);
        if constexpr (FILTER) {
            // The cells of the leaf in their order (query.py:77-85).  Every lane first moves on to its next cell that the
            // point is not surely outside of (geometry.cuh: point_surely_outside -- comparisons on the vertices just read),
            // then the lanes of the warp run the point-in-polygon test together: the points of a warp are neighbours, so
            // they share the leaf but not the cell, and testing cell after cell runs the whole test once per cell with part
            // of the lanes.
            for (int k = 0; k < leaf.y;) {
                Poly<MAXV> poly;
                int candidate = -1;
                while (k < leaf.y) {
                    const int bbox_index = leaf_element(leaf, t.bb_indices, k++);
                    pr.cell();
                    load_tree_polygon<MAXV>(t, bbox_index, poly);
                    if (!point_surely_outside(p, poly, margin, x_sides)) {
                        candidate = bbox_index;
                        break;
                    }
                }
                // (the empty statement keeps the compiler from sending every exit of the loop above straight to its own copy
                // of the test below: the lanes must meet again here, after the loop)
                asm volatile("" : "+r"(candidate));
                if (candidate >= 0 && point_in_polygon_or_on_edge(p, poly, tolerance)) return candidate;
            }
        } else {
            // the reference's loop as it stands: with two cells per leaf (the default) the bounding test above saves
            // instructions but not time, and its live values cost the 3- and 4-vertex kernels their spill-free 64 registers
            int found = -1;
            if (for_each_leaf_element(leaf, t.bb_indices, [&](int bbox_index) {
                    Poly<MAXV> poly;
                    pr.cell();
                    load_tree_polygon<MAXV>(t, bbox_index, poly);
                    if (!point_in_polygon_or_on_edge(p, poly, tolerance)) return false;
                    found = bbox_index;
                    return true;
                }))
                return found;
        }
        if (stack.empty()) return -1;
        enter(stack.pop(t));
    }
}

// ---- locate_point_on_edge, query.py:121-165 ---------------------------------------------------------------
template <bool DEEP = false>
CT_DEV int locate_point_on_edge(const TreeView &t, P2 p, double tolerance) {
    CT_STACK(stack, DEEP);
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
                stack.push(t, right_first ? left_handle : right_handle);
                cursor_descend(c, base, right_first ? right_handle : left_handle);
            } else if (left) {
                cursor_descend(c, base, left_handle);
            } else if (right) {
                cursor_descend(c, base, right_handle);
            } else {
                if (stack.empty()) return -1;
                cursor_enter(c, base, stack.pop(t));
            }
        }
        const int4 leaf = cursor_leaf(c);
        int found = -1;
        if (for_each_leaf_element(leaf, t.bb_indices, [&](int bbox_index) {
                double2 v0 = __ldg(t.elem_xy + 2 * (int64_t)bbox_index);
                double2 v1 = __ldg(t.elem_xy + 2 * (int64_t)bbox_index + 1);
                if (!point_on_edge(p, P2{v0.x, v0.y}, P2{v1.x, v1.y}, tolerance)) return false;
                found = bbox_index;
                return true;
            }))
            return found;
        if (stack.empty()) return -1;
        cursor_enter(c, base, stack.pop(t));
    }
}

// ---- locate_box, query.py:177-244 -------------------------------------------------------------------------
// Emit(bbox_index) is called for every leaf cell whose bounding box strictly overlaps `box`, in the
// reference's order (right subtree first: it pushes left then right and pops right).  Returns the count.
// A box visits ~90 nodes and the lanes of a warp are at different kinds of node most of the time (ncu: 9 of 32
// lanes active per instruction), so the step is written with as few divergent paths as possible: one for a leaf,
// one for an inner node whose outcome (left / right / both / neither) is turned into selects and a predicated push,
// and a common tail that pops if needed and always loads header + slot of the next node.
template <bool DEEP, typename Emit>
CT_DEV int locate_box(const TreeView &t, const Box4 &box, Emit emit) {
    Box4 tree_bbox{t.bbox[0], t.bbox[1], t.bbox[2], t.bbox[3]};
    if (!boxes_intersect(box, tree_bbox)) return 0;
    CT_STACK(stack, DEEP);
    int count = 0;
    const char *base = reinterpret_cast<const char *>(t.treelets);
    Cursor c;
    cursor_enter(c, base, ROOT_HANDLE);
    while (true) {
        uint32_t next = 0;
        bool pop;
        if (cursor_is_leaf(c)) {
            const int4 leaf = cursor_leaf(c);
            for_each_leaf_element(leaf, t.bb_indices, [&](int bbox_index) {
                Box4 leaf_box = load_box(t.bb_coords, bbox_index);
                if (boxes_intersect(box, leaf_box)) {
                    emit(count, bbox_index);
                    count++;
                }
                return false;
            });
            pop = true;
        } else {
            const bool dim = cursor_dim(c);
            const double bmin = dim ? box.ymin : box.xmin;
            const double bmax = dim ? box.ymax : box.xmax;
            const bool left = bmin <= c.plane.x;
            const bool right = bmax >= c.plane.y;
            uint32_t left_handle, right_handle;
            cursor_children(c, left_handle, right_handle);
            if (left && right) stack.push(t, left_handle);
            next = right ? right_handle : left_handle;
            pop = !(left || right);
        }
        if (pop) {
            if (stack.empty()) return count;
            next = stack.pop(t);
        }
        cursor_enter(c, base, next);
    }
}

// ---- per-candidate tests of locate_edge -------------------------------------------------------------------
// compute_edge_face_intersect, query.py:309-328, in its two halves: the Cohen-Sutherland clip against the cell's
// bounding box decides whether the cell is looked at at all (its clipped points are not used further) ...
CT_DEV bool edge_face_prefilter(const TreeView &t, int bbox_index, P2 a, P2 b) {
    Box4 box = load_box(t.bb_coords, bbox_index);
    return cohen_sutherland_line_meets_box(a, b, box);
}
// ... and the Cyrus-Beck clip against the cell's polygon gives the intersection
template <int MAXV>
CT_DEV bool edge_face_clip(const TreeView &t, int bbox_index, P2 a, P2 b, P2 &c, P2 &d) {
    Box4 box = load_box(t.bb_coords, bbox_index);
    Poly<MAXV> polygon;
    load_tree_polygon<MAXV>(t, bbox_index, polygon);
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
template <int MAXV, bool DEEP, typename Emit>
CT_DEV int locate_edge(const TreeView &t, P2 a, P2 b, Emit emit) {
    {
        Box4 tree_bbox{t.bbox[0], t.bbox[1], t.bbox[2], t.bbox[3]};
        if (!cohen_sutherland_line_meets_box(a, b, tree_bbox)) return 0;
    }
    P2 V = to_vector(a, b);
    CT_STACK(stack, DEEP);
    int count = 0;
    const char *base = reinterpret_cast<const char *>(t.treelets);
    Cursor cur;
    cursor_enter(cur, base, ROOT_HANDLE);
    while (true) {
        uint32_t next = 0;
        bool pop;
        if (cursor_is_leaf(cur)) {
            const int4 leaf = cursor_leaf(cur);
            for_each_leaf_element(leaf, t.bb_indices, [&](int bbox_index) {
                P2 c, d;
                if (edge_cell_intersect<MAXV>(t, bbox_index, a, b, c, d)) {
                    emit(count, bbox_index, c, d);
                    count++;
                }
                return false;
            });
            pop = true;
        } else {
            bool left, right;
            edge_plane_test(cur, a, b, V, left, right);
            uint32_t left_handle, right_handle;
            cursor_children(cur, left_handle, right_handle);
            if (left && right) stack.push(t, left_handle);
            next = right ? right_handle : left_handle;
            pop = !(left || right);
        }
        if (pop) {
            if (stack.empty()) return count;
            next = stack.pop(t);
        }
        cursor_enter(cur, base, next);
    }
}

}  // namespace ct
