// geometry.cuh -- fp64 device functions for the per-candidate geometry of the query path.
//
// Each function states the reference function it must agree with bit-for-bit (file:line relative to the
// reference's numba_celltree/ package).  Rules that keep the bits equal:
//   * expressions are evaluated in the reference's source order, no FMA contraction (-fmad=false);
//   * Numba's builtin min(a, b) is select(b < a, b, a) and max(a, b) is select(b > a, b, a): ties and
//     NaNs keep the FIRST argument -> nb_min / nb_max, never fmin / fmax;
//   * predicates keep their order and short-circuiting (a division by zero behind a false guard must not
//     change a result; IEEE inf/NaN propagate exactly as on the CPU).
//
// Polygons live in per-thread arrays sized by the template bound MAXV (3, 4, 8 or 32 vertices); for the
// small bounds every loop is unrolled so the vertices stay in registers.
#pragma once

#include "common.cuh"

namespace ct {

struct P2 {
    double x, y;
};

template <int MAXV>
struct Poly {
    double x[MAXV];
    double y[MAXV];
    int n;
};

#define CT_DEV __device__ __forceinline__

CT_DEV double nb_min(double a, double b) { return b < a ? b : a; }
CT_DEV double nb_max(double a, double b) { return b > a ? b : a; }

CT_DEV P2 to_vector(P2 a, P2 b) { return P2{b.x - a.x, b.y - a.y}; }                // geometry_utils.py:23-25
CT_DEV double cross_product(P2 u, P2 v) { return u.x * v.y - u.y * v.x; }             // :57-59
CT_DEV double dot_product(P2 u, P2 v) { return u.x * v.x + u.y * v.y; }               // :62-64
CT_DEV double length_squared(P2 v) { return v.x * v.x + v.y * v.y; }                  // :67-69
CT_DEV P2 to_point(double t, P2 a, P2 V) { return P2{a.x + t * V.x, a.y + t * V.y}; }  // :52-54

// Indexed read that stays in registers for the small bounds (select chain instead of local memory).
template <int MAXV>
CT_DEV P2 pget(const Poly<MAXV> &p, int idx) {
    if constexpr (MAXV <= 8) {
        P2 r{p.x[0], p.y[0]};
#pragma unroll
        for (int k = 1; k < MAXV; k++)
            if (k == idx) {
                r.x = p.x[k];
                r.y = p.y[k];
            }
        return r;
    } else {
        return P2{p.x[idx], p.y[idx]};
    }
}

// copy_vertices_into (geometry_utils.py:501-510) with polygon_length (:72-79): a face row is scanned for
// the first -1 from column 3 on; at least 3 vertices are always read.
// `xy` holds the vertex coordinates per element row, (n, M) double2 (tree: ct_tree.elem_xy, built once from
// faces + vertices), so the id row (for the length) and the coordinates are two INDEPENDENT loads.
template <int MAXV>
CT_DEV void load_polygon(const int32_t *__restrict__ elements, int M, int64_t elem, const double2 *__restrict__ xy,
                         Poly<MAXV> &poly) {
    const int32_t *row = elements + elem * (int64_t)M;
    const double2 *c = xy + elem * (int64_t)M;
    int n = M < MAXV ? M : MAXV;
    if constexpr (MAXV <= 4) {
        // all coordinates are requested before the length is known (the row of xy is complete, padding is (0, 0)):
        // the id row and the coordinates arrive together instead of one after the other
        double2 v[MAXV];
#pragma unroll
        for (int k = 0; k < MAXV; k++) v[k] = (k < M) ? __ldg(c + k) : make_double2(0.0, 0.0);
        if constexpr (MAXV == 4) {
            if (M == 4) {
                int4 r = __ldg(reinterpret_cast<const int4 *>(row));
                if (r.w == -1) n = 3;
            }
        }
        poly.n = n;
#pragma unroll
        for (int k = 0; k < MAXV; k++) {
            poly.x[k] = v[k].x;
            poly.y[k] = v[k].y;
        }
    } else {
#pragma unroll
        for (int k = MAXV - 1; k >= 3; k--)
            if (k < M && __ldg(row + k) == -1) n = k;
        poly.n = n;
#pragma unroll
        for (int k = 0; k < MAXV; k++) {
            if (k < n) {
                double2 v = __ldg(c + k);
                poly.x[k] = v.x;
                poly.y[k] = v.y;
            }
        }
    }
}

// A cell of the tree.  Rows of up to four vertices: the coordinates alone are read (one contiguous access); a fourth
// vertex that is the padding marker means a triangle in a quad row (polygon_length, geometry_utils.py:72-79, scans the
// id row for the first -1 from column 3 on -- the marker sits exactly where the id row holds -1).
template <int MAXV>
CT_DEV void load_tree_polygon(const TreeView &t, int64_t elem, Poly<MAXV> &poly) {
    if constexpr (MAXV <= 4) {
        if (!t.length_from_rows) {
            constexpr int M = MAXV;  // every launcher picks the 3- and 4-vertex kernels for exactly that row width
            const double2 *c = t.elem_xy + elem * (int64_t)M;
            double2 v[MAXV];
#pragma unroll
            for (int k = 0; k < MAXV; k++) v[k] = __ldg(c + k);
            int n = MAXV;
            if constexpr (MAXV == 4) {
                if ((unsigned long long)__double_as_longlong(v[3].x) == PAD_VERTEX_BITS) n = 3;
            }
            poly.n = n;
#pragma unroll
            for (int k = 0; k < MAXV; k++) {
                poly.x[k] = v[k].x;
                poly.y[k] = v[k].y;
            }
            return;
        }
    }
    load_polygon<MAXV>(t.elements, t.M, elem, t.elem_xy, poly);
}

// The same for a polygon given by vertex ids into a vertex array (query faces of ct_locate_faces).
template <int MAXV>
CT_DEV void gather_polygon(const int32_t *__restrict__ elements, int M, int64_t elem, const double2 *__restrict__ vertices,
                           Poly<MAXV> &poly) {
    int idx[MAXV];
    const int32_t *row = elements + elem * (int64_t)M;
#pragma unroll
    for (int k = 0; k < MAXV; k++) idx[k] = (k < M) ? __ldg(row + k) : -1;
    int n = M < MAXV ? M : MAXV;
#pragma unroll
    for (int k = MAXV - 1; k >= 3; k--)
        if (k < M && idx[k] == -1) n = k;
    poly.n = n;
#pragma unroll
    for (int k = 0; k < MAXV; k++) {
        if (k < n) {
            double2 v = __ldg(vertices + idx[k]);
            poly.x[k] = v.x;
            poly.y[k] = v.y;
        }
    }
}

CT_DEV bool within_perpendicular_distance(double UxV, P2 U, double tolerance) {  // geometry_utils.py:150-166
    return (UxV * UxV) < ((tolerance * length_squared(U)) * tolerance);
}

CT_DEV bool in_bounds(P2 p, P2 a, P2 b) {  // geometry_utils.py:169-192
    double xmin = nb_min(a.x, b.x);
    double xmax = nb_max(a.x, b.x);
    double ymin = nb_min(a.y, b.y);
    double ymax = nb_max(a.y, b.y);
    double dx = xmax - xmin;
    double dy = ymax - ymin;
    bool use_x_bound = fabs(dx) >= fabs(dy);
    return (use_x_bound && ((p.x >= xmin) && (p.x <= xmax))) || (!use_x_bound && ((p.y >= ymin) && (p.y <= ymax)));
}

// a / b, IEEE-exact.  A zero numerator always takes the software division's slow path (some 70 instructions);
// on rectilinear meshes every crossing test divides (v1.x - v0.x) * ... = 0 for the vertical edges, so that case
// is answered directly: (+-0) / (nonzero, non-NaN b) = zero with the sign of a XOR the sign of b.
CT_DEV double div_exact(double a, double b) {
    if (a == 0.0 && b != 0.0 && b == b)
        return __longlong_as_double((__double_as_longlong(a) ^ __double_as_longlong(b)) & (long long)0x8000000000000000ULL);
    return a / b;
}

// One edge v0 -> v1 of point_in_polygon_or_on_edge (geometry_utils.py:203-220): U = v0 - p is carried from the previous
// edge, V = v1 - p becomes the next U.  Returns true when the point is accepted on this edge; flips `c` on a crossing.
CT_DEV bool pip_edge(P2 p, P2 v0, P2 v1, P2 U, P2 &V, double tolerance, bool &c) {
    V = to_vector(p, v1);
    const double A = cross_product(U, V);
    const P2 W = to_vector(v0, v1);
    if (within_perpendicular_distance(A, W, tolerance) && in_bounds(p, v0, v1)) return true;
    if (((v0.y > p.y) != (v1.y > p.y)) && (p.x < (div_exact((v1.x - v0.x) * (p.y - v0.y), (v1.y - v0.y)) + v0.x))) c = !c;
    return false;
}

// geometry_utils.py:195-223 -- crossing-number test with tolerance-based acceptance on the boundary.
//
// Two facts about the reference's loop let the 3- and 4-vertex forms below run without its two data-dependent parts:
//  * its result is (some edge accepts the point) OR (the number of crossings is odd): both are independent of the order in
//    which the edges are visited, and every edge's arithmetic only involves its own two vertices and the point;
//  * an edge of zero length (`if v1 == v0: continue`) changes nothing even when it is not skipped: U and V are then the same
//    vector (up to the sign of a zero), A = U.x * V.y - U.y * V.x is a zero, W is a zero vector, so the acceptance test reads
//    0 < (tolerance * 0) * tolerance -- false, also for an infinite or NaN tolerance -- and the crossing test reads
//    (v.y > p.y) != (v.y > p.y), false; where v1 and v0 differ in the sign of a zero, later comparisons cannot tell.
// So a triangle stored in a four-vertex row is walked as the quad (v0, v1, v2, v0): its edges are the triangle's three
// plus the zero-length edge v0 -> v0, and no vertex has to be picked by a run-time index.
template <int MAXV>
CT_DEV bool point_in_polygon_or_on_edge(P2 p, const Poly<MAXV> &poly, double tolerance) {
    if constexpr (MAXV == 3 || MAXV == 4) {
        const P2 a{poly.x[0], poly.y[0]}, b{poly.x[1], poly.y[1]}, d{poly.x[2], poly.y[2]};
        P2 last = d;
        if constexpr (MAXV == 4) {
            if (poly.n == 4) last = P2{poly.x[3], poly.y[3]};
            else last = a;  // triangle in a quad row: closing edge d -> a, then the zero-length edge a -> a
        }
        bool c = false;
        P2 U = to_vector(p, last), V;
        if (pip_edge(p, last, a, U, V, tolerance, c)) return true;
        U = V;
        if (pip_edge(p, a, b, U, V, tolerance, c)) return true;
        U = V;
        if (pip_edge(p, b, d, U, V, tolerance, c)) return true;
        if constexpr (MAXV == 4) {
            U = V;
            if (pip_edge(p, d, last, U, V, tolerance, c)) return true;
        }
        return c;
    } else {
        const int length = poly.n;
        P2 v0 = pget(poly, length - 1);
        P2 U = to_vector(p, v0);
        bool c = false;
#pragma unroll
        for (int i = 0; i < MAXV; i++) {
            if (i >= length) break;
            P2 v1{poly.x[i], poly.y[i]};
            if (v1.x == v0.x && v1.y == v0.y) continue;
            P2 V;
            if (pip_edge(p, v0, v1, U, V, tolerance, c)) return true;
            v0 = v1;
            U = V;
        }
        return c;
    }
}

// True only where point_in_polygon_or_on_edge above is CERTAIN to answer false: the point lies beyond the polygon's bounding
// box by more than a margin.  The reference has no such test (query.py:77-85 runs the full test on every cell of a leaf);
// it is here so that the lanes of a warp first pick their candidate cell -- neighbouring points sit in different cells of the
// same leaf -- and then run the expensive test together, once, instead of one after the other on every cell.
//
// Why the answer is certain.  Let the point be beyond every vertex in one coordinate, say p.x - v.x > m for all vertices
// (the other three sides alike), m = 8 |tolerance| + 1e-12 * (sum of |v.x|) + 2e-150.
//  * Acceptance on an edge needs in_bounds: an edge bounded in x fails it exactly (comparisons only).  For an edge bounded
//    in y (|dy| > |dx|) with p.y inside its range, U.y and V.y have opposite signs and U.x, V.x the same sign, so the two
//    products of A = U.x V.y - U.y V.x have the same sign -- no cancellation: |A| >= m (|U.y| + |V.y|) = m |dy| >= m |W| / sqrt 2,
//    A^2 >= 32 tolerance^2 |W|^2, while the right-hand side tolerance * |W|^2 * tolerance is at most 8 times its exact
//    value even where |W|^2 or the products fall into the denormal range (each of three roundings at most doubles a denormal):
//    not accepted.
//  * The crossing test: beyond in y there is no straddling edge (exact comparisons).  Beyond in x, every straddling edge
//    computes its intersection abscissa (dx * (p.y - v0.y)) / dy + v0.x within 16 ulp-sized errors of the edge's x-range --
//    1e-12 * sum |v.x| is 250 times that -- unless the product underflows, where the error is at most |dx| (the quotient
//    lies between 0 and 2 dx): |dx| < 1e-150 is covered by the 2e-150 of the margin, and |dx| >= 1e-150 needs
//    0 < |p.y - v0.y| < 2.2e-158 to underflow, which outside_x_sides_allowed() below excludes once per point.  So all
//    straddling edges answer alike, and their number is even (the booleans v.y > p.y around a closed polygon change an even
//    number of times): c stays false.  Huge coordinates, whose products could overflow, switch the x sides off.
// Comparisons with NaN are false, so a NaN anywhere (coordinates, point, tolerance) never rejects.
template <int MAXV>
CT_DEV bool point_surely_outside(P2 p, const Poly<MAXV> &poly, double margin, bool x_sides_allowed) {
    double sx = 0.0, sy = 0.0;
    bool xhi = true, xlo = true, yhi = true, ylo = true;
    auto vertex = [&](double vx, double vy) {
        sx += fabs(vx);
        sy += fabs(vy);
    };
    auto side = [&](double vx, double vy, double mx, double my) {
        const double dx = p.x - vx, dy = p.y - vy;
        xhi = xhi && dx > mx;
        xlo = xlo && dx < -mx;
        yhi = yhi && dy > my;
        ylo = ylo && dy < -my;
    };
    if constexpr (MAXV == 3 || MAXV == 4) {
        // a triangle in a four-vertex row counts its first vertex twice (the fourth slot holds the padding marker)
        const bool four = MAXV == 4 && poly.n == 4;
        const double x3 = four ? poly.x[MAXV - 1] : poly.x[0], y3 = four ? poly.y[MAXV - 1] : poly.y[0];
#pragma unroll
        for (int k = 0; k < 3; k++) vertex(poly.x[k], poly.y[k]);
        vertex(x3, y3);
        const double mx = fma(sx, 1e-12, margin), my = fma(sy, 1e-12, margin);
#pragma unroll
        for (int k = 0; k < 3; k++) side(poly.x[k], poly.y[k], mx, my);
        side(x3, y3, mx, my);
    } else {
#pragma unroll
        for (int k = 0; k < MAXV; k++)
            if (k < poly.n) vertex(poly.x[k], poly.y[k]);
        const double mx = fma(sx, 1e-12, margin), my = fma(sy, 1e-12, margin);
#pragma unroll
        for (int k = 0; k < MAXV; k++)
            if (k < poly.n) side(poly.x[k], poly.y[k], mx, my);
    }
    return yhi || ylo || ((xhi || xlo) && x_sides_allowed && (sx + sy) < 1e100);
}
// Whether the x sides of point_surely_outside may be used for this point: a point whose |y| is at least 1e-140 differs from
// any other y coordinate by zero or by more than 5e-157 (the difference of two doubles is a multiple of the smaller one's
// unit in the last place), so with |dx| >= 1e-150 the product dx * (p.y - v0.y) of the crossing test cannot underflow.
CT_DEV bool outside_x_sides_allowed(P2 p) { return fabs(p.y) >= 1e-140; }
CT_DEV double outside_margin(double tolerance) { return fma(8.0, fabs(tolerance), 2e-150); }

CT_DEV bool point_on_edge(P2 p, P2 v0, P2 v1, double tolerance) {  // geometry_utils.py:226-238
    if (v1.x == v0.x && v1.y == v0.y) return false;
    P2 U = to_vector(p, v0);
    P2 V = to_vector(p, v1);
    P2 W = to_vector(v0, v1);
    double A = cross_product(U, V);
    return within_perpendicular_distance(A, W, tolerance) && in_bounds(p, v0, v1);
}

struct Box4 {
    double xmin, xmax, ymin, ymax;
};

CT_DEV Box4 load_box(const double *__restrict__ rows, int64_t i) {
    const double2 *p = reinterpret_cast<const double2 *>(rows + 4 * i);
    double2 a = __ldg(p), b = __ldg(p + 1);
    return Box4{a.x, a.y, b.x, b.y};
}

CT_DEV bool boxes_intersect(const Box4 &a, const Box4 &b) {  // geometry_utils.py:292-300
    return a.xmin < b.xmax && b.xmin < a.xmax && a.ymin < b.ymax && b.ymin < a.ymax;
}

// ---- segment / segment (EdgeCellTree2d.intersect_edges) ------------------------------------------------
CT_DEV bool left_of(P2 a, P2 p, P2 U) { return U.x * (a.y - p.y) > U.y * (a.x - p.x); }  // geometry_utils.py:318-322

CT_DEV bool has_overlap(double a, double b, double p, double q, double tolerance) {  // :325-329
    return ((nb_min(a, b) - nb_max(p, q)) < tolerance) && ((nb_min(p, q) - nb_max(a, b)) < tolerance);
}

CT_DEV P2 intersection_location_point(P2 V, P2 U, P2 a, P2 p, double tolerance) {  // :332-346
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    double denom = cross_product(V, U);
    if (within_perpendicular_distance(denom, V, tolerance)) return P2{nan, nan};
    P2 R = to_vector(a, p);
    double t = cross_product(R, U) / denom;
    return P2{a.x + t * V.x, a.y + t * V.y};
}

CT_DEV P2 midpoint_collinear_lines(P2 a, P2 b, P2 p, P2 q) {  // :349-374
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    if (a.x > b.x || (a.x == b.x && a.y > b.y)) { P2 t = a; a = b; b = t; }
    if (p.x > q.x || (p.x == q.x && p.y > q.y)) { P2 t = p; p = q; q = t; }
    double overlap_start_x = nb_max(a.x, p.x);
    double overlap_start_y = nb_max(a.y, p.y);
    double overlap_end_x = nb_min(b.x, q.x);
    double overlap_end_y = nb_min(b.y, q.y);
    if (overlap_start_x > overlap_end_x || overlap_start_y > overlap_end_y) return P2{nan, nan};
    return P2{0.5 * (overlap_start_x + overlap_end_x), 0.5 * (overlap_start_y + overlap_end_y)};
}

CT_DEV bool lines_intersect(P2 a, P2 b, P2 p, P2 q, P2 &out) {  // :377-418
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    out = P2{nan, nan};
    P2 V = to_vector(a, b);
    P2 U = to_vector(p, q);
    if ((U.x == 0 && U.y == 0) || (V.x == 0 && V.y == 0)) return false;
    if ((U.x == 0) && (V.x == 0) && a.x != p.x) return false;
    if ((U.y == 0) && (V.y == 0) && a.y != p.y) return false;
    double tolerance = nb_max(MIN_TOLERANCE, TOLERANCE_FACTOR * nb_max(fabs(U.x), fabs(U.y)));
    if ((!has_overlap(a.x, b.x, p.x, q.x, tolerance)) || (!has_overlap(a.y, b.y, p.y, q.y, tolerance))) return false;
    if ((left_of(a, p, U) != left_of(b, p, U)) && (left_of(p, a, V) != left_of(q, a, V))) {
        out = intersection_location_point(V, U, a, p, tolerance);
        return true;
    }
    P2 R = to_vector(a, p);
    P2 S = to_vector(a, q);
    if (within_perpendicular_distance(cross_product(V, R), V, tolerance) &&
        within_perpendicular_distance(cross_product(V, S), V, tolerance)) {
        out = midpoint_collinear_lines(a, b, p, q);
        return true;
    }
    return false;
}

// ---- Cohen-Sutherland segment / box: algorithms/cohen_sutherland.py ----------------------------------
enum { CS_INSIDE = 0, CS_LEFT = 1, CS_RIGHT = 2, CS_LOWER = 4, CS_UPPER = 8 };

CT_DEV int get_clip(P2 a, const Box4 &box) {  // cohen_sutherland.py:18-33
    int p = CS_INSIDE;
    if (a.x < box.xmin) p |= CS_LEFT;
    else if (a.x > box.xmax) p |= CS_RIGHT;
    if (a.y < box.ymin) p |= CS_LOWER;
    else if (a.y > box.ymax) p |= CS_UPPER;
    return p;
}

// cohen_sutherland.py:36-101.  Returns 1 (c, d = clipped segment), 0 (no intersection; c, d = NaN).
// The reference's "Undefined clipping state" branch is unreachable (a non-zero outcode has one of the
// four bits set), so there is no error return.
// WANT_POINTS = false: only whether the segment meets the box (the prefilter of the segment queries, query.py:316-319,
// throws the clipped points away).  The function is out of line, so its outputs go through memory: without them the
// segment walk saves eight 8-byte local stores per candidate (1.4 G of its 1.8 G local store sectors on C4).
template <bool WANT_POINTS>
static __device__ __noinline__ int cohen_sutherland_clip(P2 a, P2 b, Box4 box, P2 *c, P2 *d) {
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    if constexpr (WANT_POINTS) {
        *c = P2{nan, nan};
        *d = P2{nan, nan};
    }
    double dx = b.x - a.x;
    double dy = b.y - a.y;
    if (dx == 0.0 && dy == 0.0) return 0;
    int k1 = get_clip(a, box);
    int k2 = get_clip(b, box);
    while ((k1 | k2) != CS_INSIDE) {
        if ((k1 & k2) != 0) return 0;
        int opt = k1 ? k1 : k2;
        // The four cases of cohen_sutherland.py:71-84 (upper, lower, right, left, in that priority) are one expression
        // with the operands exchanged: moved = along + (slope * (bound - across)) / run.  Selecting the operands and
        // dividing once keeps the lanes of a warp together (one division sequence instead of four divergent ones); the
        // operations and their order are the reference's, so the bits are.
        const bool horizontal = (opt & (CS_UPPER | CS_LOWER)) != 0;  // clip against y = bound, else against x = bound
        const double bound = (opt & CS_UPPER) ? box.ymax : ((opt & CS_LOWER) ? box.ymin : ((opt & CS_RIGHT) ? box.xmax : box.xmin));
        const double moved = (horizontal ? a.x : a.y) + ((horizontal ? dx : dy) * (bound - (horizontal ? a.y : a.x))) / (horizontal ? dy : dx);
        const double x = horizontal ? moved : bound, y = horizontal ? bound : moved;
        if (opt == k1) { a = P2{x, y}; k1 = get_clip(a, box); }
        else if (opt == k2) { b = P2{x, y}; k2 = get_clip(b, box); }
        dx = b.x - a.x;
        dy = b.y - a.y;
        if (dx == 0.0 && dy == 0.0) return 0;
    }
    if constexpr (WANT_POINTS) {
        *c = a;
        *d = b;
    }
    return 1;
}
CT_DEV int cohen_sutherland_line_box_clip(P2 a, P2 b, Box4 box, P2 &c, P2 &d) { return cohen_sutherland_clip<true>(a, b, box, &c, &d); }
CT_DEV bool cohen_sutherland_line_meets_box(P2 a, P2 b, Box4 box) { return cohen_sutherland_clip<false>(a, b, box, nullptr, nullptr) != 0; }

// ---- Cyrus-Beck / Skala segment / convex polygon: algorithms/cyrus_beck.py -----------------------------
CT_DEV bool cb_compute_intersection(P2 a, P2 s, P2 v0, P2 v1, double &t) {  // cyrus_beck.py:35-52
    P2 si = to_vector(a, v0);
    P2 n{-(v1.y - v0.y), (v1.x - v0.x)};
    double n_si = dot_product(n, si);
    double k = dot_product(n, s);
    t = n_si / k;
    return n_si > 0;
}

CT_DEV bool cb_overlap(double ta, double tb, double t0, double t1) {  // :75-82
    if (ta > tb) { double t = ta; ta = tb; tb = t; }
    if (t0 > t1) { double t = t0; t0 = t1; t1 = t; }
    double vector_overlap = nb_max(0.0, nb_min(tb, t1) - nb_max(ta, t0));
    return vector_overlap > 0.0;
}

CT_DEV bool cb_aligned(P2 U, P2 V) {  // :85-100
    if ((U.x == 0 && U.y == 0) || (V.x == 0 && V.y == 0)) return true;
    if (U.x != 0 && V.x != 0) return (U.x > 0) == (V.x > 0);
    if (U.y != 0 && V.y != 0) return (U.y > 0) == (V.y > 0);
    return false;
}

CT_DEV bool cb_collinear_case(P2 a, P2 b, P2 v0, P2 v1, P2 &c, P2 &d) {  // :103-139
    P2 _b{b.x - a.x, b.y - a.y};
    P2 _v0{v0.x - a.x, v0.y - a.y};
    P2 _v1{v1.x - a.x, v1.y - a.y};
    P2 U = _b;
    P2 V = to_vector(_v0, _v1);
    if (!cb_aligned(U, V)) {
        P2 t = v0; v0 = v1; v1 = t;
        t = _v0; _v0 = _v1; _v1 = t;
    }
    P2 n{-_b.y, _b.x};
    double ta = 0.0;
    double tb = cross_product(n, _b);
    double t0 = cross_product(n, _v0);
    double t1 = cross_product(n, _v1);
    if (!cb_overlap(ta, tb, t0, t1)) return false;
    c = (t0 < ta) ? v0 : a;
    d = (t1 > tb) ? v1 : b;
    return true;
}

// cyrus_beck.py:143-241.  Polygon must be counter-clockwise.  c, d are NaN when false is returned.
template <int MAXV>
__device__ __noinline__ bool cyrus_beck_line_polygon_clip(P2 a, P2 b, const Poly<MAXV> &poly, double tolerance, P2 &c, P2 &d) {
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    c = P2{nan, nan};
    d = P2{nan, nan};
    const int length = poly.n;
    P2 s = to_vector(a, b);
    if (s.x == 0 && s.y == 0) return false;
    bool a_inside = point_in_polygon_or_on_edge(a, poly, tolerance);
    bool b_inside = point_in_polygon_or_on_edge(b, poly, tolerance);
    if (a_inside && b_inside) {
        c = a;
        d = b;
        return true;
    }
    int i0 = -1, i1 = -1, i = 0, k = 0;
    P2 v0{poly.x[0], poly.y[0]};
    double ksi = cross_product(to_vector(a, v0), s);
    while (i < length && k < 2) {
        int inext = (i + 1 == length) ? 0 : i + 1;
        P2 v1 = pget(poly, inext);
        double eta = cross_product(to_vector(a, v1), s);
        if ((ksi < 0.0) != (eta < 0.0)) {
            if (k == 0) i0 = i;
            else i1 = i;
            k += 1;
        } else if ((ksi == 0.0) && (eta == 0.0)) {
            bool r = cb_collinear_case(a, b, v0, v1, c, d);
            if (!r) { c = P2{nan, nan}; d = P2{nan, nan}; }
            return r;
        }
        ksi = eta;
        v0 = v1;
        i += 1;
    }
    if (k == 0) return false;

    // intersections(), cyrus_beck.py:55-72.  With a single crossing i1 == -1, which the reference uses as a
    // Python negative index: edge (last vertex -> first vertex).
    double t0, t1;
    {
        int j1 = (i1 < 0) ? i1 + length : i1;
        int j0n = (i0 + 1 == length) ? 0 : i0 + 1;
        int j1n = (i1 + 1) % length;
        double ta, tb;
        (void)cb_compute_intersection(a, s, pget(poly, i0), pget(poly, j0n), ta);
        bool enters1 = cb_compute_intersection(a, s, pget(poly, j1), pget(poly, j1n), tb);
        if (enters1) { t0 = tb; t1 = ta; }
        else { t0 = ta; t1 = tb; }
    }
    if (t0 == t1) {
        if (a_inside && t1 != 0.0) t0 = 0.0;
        else if (b_inside && t0 != 1.0) t1 = 1.0;
        else return false;
    }
    if (t1 < t0) { double t = t0; t0 = t1; t1 = t; }
    bool valid0 = t0 >= 0 && t0 < 1;
    bool valid1 = t1 > 0 && t1 <= 1;
    if (valid0 && valid1) { c = to_point(t0, a, s); d = to_point(t1, a, s); return true; }
    else if (valid0) { c = to_point(t0, a, s); d = b; return true; }
    else if (valid1) { c = a; d = to_point(t1, a, s); return true; }
    return false;
}

// ---- Sutherland-Hodgman convex clip + area: algorithms/sutherland_hodgman.py -------------------------
CT_DEV bool sh_inside(P2 p, P2 r, P2 U) { return U.x * (p.y - r.y) > U.y * (p.x - r.x); }  // :56-60

CT_DEV bool sh_intersection(P2 a, P2 V, P2 r, P2 N, P2 &out) {  // :63-74
    P2 W{r.x - a.x, r.y - a.y};
    double nw = dot_product(N, W);
    double nv = dot_product(N, V);
    if (nv != 0) {
        double t = nw / nv;
        out = P2{a.x + t * V.x, a.y + t * V.y};
        return true;
    }
    return false;
}

// Working storage of the clip: two polygons of CAP vertices that swap roles after every clip edge.
// Small bounds live in shared memory, vertex-major with the block's threads innermost ([2][CAP][threads] double2:
// conflict-free 16-byte accesses); as per-thread arrays they would be indexed dynamically and end up in local
// memory, whose traffic through L1/L2 was what bound the pair kernels (234 GB of L2 sectors per 126 M pairs).
template <int CAP>
struct SharedClipWork {
    double2 *base;  // this thread's first element
    int stride;     // threads per block
    CT_DEV double2 &at(int buffer, int j) const { return base[(buffer * CAP + j) * stride]; }
};
template <int CAP>
struct LocalClipWork {
    double2 v[2][CAP];
    CT_DEV double2 &at(int buffer, int j) { return v[buffer][j]; }
};

// polygon_area, geometry_utils.py:82-95 (fan triangulation, abs of each cross product)
template <typename Work>
CT_DEV double polygon_area(Work &work, int buffer, int length) {
    double area = 0.0;
    const double2 a2 = work.at(buffer, 0), b2 = work.at(buffer, 1);
    P2 a{a2.x, a2.y};
    P2 U = to_vector(a, P2{b2.x, b2.y});
    for (int i = 2; i < length; i++) {
        const double2 c2 = work.at(buffer, i);
        P2 V = to_vector(P2{c2.x, c2.y}, a);
        area += fabs(cross_product(U, V));
        U = V;
    }
    return 0.5 * area;
}

// polygon_polygon_clip_area, sutherland_hodgman.py:84-148: clip `polygon` (subject) by every edge of
// `clipper`; zero-length clipper / subject edges are skipped; early 0.0 when fewer than 3 vertices remain.
// The reference sizes its working polygons 2 * MAX_N_VERTEX = 64 and copies the output back into the subject after
// every clip edge; here two buffers of CAP vertices swap roles.  A convex subject gains at most one vertex per clip
// edge, so CAP = MAXA + MAXB holds it; a concave or self-intersecting cell can gain more (up to one per subject edge
// and clip edge): `overflow` is then set instead of dropping a vertex, and the caller repeats the pair with the
// reference's capacity.
template <int MAXA, int MAXB, int CAP, typename Work>
CT_DEV double polygon_polygon_clip_area(const Poly<MAXA> &polygon, const Poly<MAXB> &clipper, Work &work, bool &overflow) {
    int n_output = polygon.n;
    const int n_clip = clipper.n;
    int out = 0;  // buffer that holds the current output polygon
    overflow = false;
#pragma unroll
    for (int i = 0; i < MAXA; i++)
        if (i < n_output) work.at(out, i) = make_double2(polygon.x[i], polygon.y[i]);

    P2 r = pget(clipper, n_clip - 1);
#pragma unroll
    for (int i = 0; i < MAXB; i++) {
        if (i >= n_clip) break;
        P2 s{clipper.x[i], clipper.y[i]};
        P2 U{s.x - r.x, s.y - r.y};
        if (U.x == 0 && U.y == 0) continue;
        P2 N{-U.y, U.x};
        const int length = n_output;
        const int in = out;  // the previous output is the subject of this clip edge
        out ^= 1;
        n_output = 0;
        auto push = [&](P2 v) {
            if (n_output < CAP) {
                work.at(out, n_output) = make_double2(v.x, v.y);
                n_output++;
            } else {
                overflow = true;
            }
        };
        const double2 last = work.at(in, length - 1);
        P2 a{last.x, last.y};
        bool a_inside = sh_inside(a, r, U);
        for (int j = 0; j < length; j++) {
            const double2 b2 = work.at(in, j);
            P2 b{b2.x, b2.y};
            P2 V{b.x - a.x, b.y - a.y};
            if (V.x == 0 && V.y == 0) continue;
            bool b_inside = sh_inside(b, r, U);
            // The reference's four cases (sutherland_hodgman.py:118-137) with ONE site for the intersection and one for each
            // kind of push: an edge that enters the clip half-plane and one that leaves it need the same intersection (a
            // division: a fifth of this function's instructions), and with a site of their own each they ran it one after the
            // other, at five active lanes.  Entering: push the intersection (if the edge is not parallel), then b.  Leaving:
            // push the intersection -- or, if parallel, count b as inside and push it.  Both inside: push b.
            const bool crossing = a_inside != b_inside;
            bool have = false;
            P2 point;
            if (crossing) have = sh_intersection(a, V, r, N, point);
            if (have) push(point);
            if (b_inside || (crossing && !have)) {
                b_inside = true;
                push(b);
            }
            a = b;
            a_inside = b_inside;
        }
        if (overflow) return 0.0;
        if (n_output < 3) return 0.0;
        r = s;
    }
    return polygon_area(work, out, n_output);
}

// the pair again with working polygons of the reference's size (out of line: almost never taken)
template <int MAXA, int MAXB>
__device__ __noinline__ double clip_area_full_capacity(const Poly<MAXA> &a, const Poly<MAXB> &b) {
    LocalClipWork<2 * MAX_N_VERTEX> work;
    bool overflow;
    return polygon_polygon_clip_area<MAXA, MAXB, 2 * MAX_N_VERTEX>(a, b, work, overflow);
}

// The per-pair kernels' entry: shared-memory working polygons for the small bounds, per-thread arrays otherwise.
// THREADS = threads per block of the calling kernel.
template <int MAXA, int MAXB, int THREADS>
CT_DEV double clip_area_of_pair(const Poly<MAXA> &a, const Poly<MAXB> &b) {
    constexpr int CAP = MAXA + MAXB;
    bool overflow;
    double area;
    if constexpr (CAP <= 8) {
        __shared__ double2 storage[2 * CAP * THREADS];
        SharedClipWork<CAP> work{storage + threadIdx.x, THREADS};
        area = polygon_polygon_clip_area<MAXA, MAXB, CAP>(a, b, work, overflow);
    } else {
        LocalClipWork<CAP> work;
        area = polygon_polygon_clip_area<MAXA, MAXB, CAP>(a, b, work, overflow);
    }
    if (overflow) area = clip_area_full_capacity<MAXA, MAXB>(a, b);
    return area;
}

// ---- separating axis test: algorithms/separating_axis.py -----------------------------------------------
template <int MAXV>
CT_DEV void extrema_projected(P2 norm, const Poly<MAXV> &polygon, double &mn, double &mx) {  // :17-27
    double min_proj = FLOAT_MAX, max_proj = FLOAT_MIN;
#pragma unroll
    for (int i = 0; i < MAXV; i++) {
        if (i >= polygon.n) break;
        double proj = dot_product(P2{polygon.x[i], polygon.y[i]}, norm);
        min_proj = nb_min(min_proj, proj);
        max_proj = nb_max(max_proj, proj);
    }
    mn = min_proj;
    mx = max_proj;
}

// separating_axes(a, b), :42-55: true when no edge normal of `a` separates the two polygons
template <int MAXA, int MAXB>
CT_DEV bool separating_axes(const Poly<MAXA> &a, const Poly<MAXB> &b) {
    P2 p = pget(a, a.n - 1);
#pragma unroll
    for (int i = 0; i < MAXA; i++) {
        if (i >= a.n) break;
        P2 q{a.x[i], a.y[i]};
        P2 norm{p.y - q.y, q.x - p.x};
        p = q;
        if (norm.x == 0.0 && norm.y == 0.0) continue;
        double mina, maxa, minb, maxb;  // is_separating_axis, :30-39
        extrema_projected(norm, a, mina, maxa);
        extrema_projected(norm, b, minb, maxb);
        if (!(maxa > minb && maxb > mina)) return false;
    }
    return true;
}

// ---- barycentric weights ---------------------------------------------------------------------------------
// algorithms/barycentric_triangle.py:28-43
// A triangle of zero area gives 1.0 / 0.0 = inf and NaN weights, in the reference as here: its compute_weights is inlined
// into the parallel loop, where Numba's division does not raise (measured: NaN rows, no exception, any batch size).
CT_DEV void triangle_weights(P2 a, P2 b, P2 c, P2 p, double &u, double &v, double &w) {
    P2 ab = to_vector(a, b);
    P2 ac = to_vector(a, c);
    P2 ap = to_vector(a, p);
    double Aa = fabs(cross_product(ab, ap));
    double Ac = fabs(cross_product(ac, ap));
    double A = fabs(cross_product(ab, ac));
    double inv_denom = 1.0 / A;
    w = inv_denom * Aa;
    v = inv_denom * Ac;
    u = 1.0 - v - w;
}

// algorithms/barycentric_wachspress.py:38-85 (+ interp_edge_case :26-35).  w[] has MAXV slots, all
// pre-zeroed by the caller; slots >= poly.n stay zero (the reference's rows are n_max_vert wide).
// `zero_division` is set where this function divides by exactly zero.  The reference's compute_weights is a separate
// @njit function with Python's error model, so it raises ZeroDivisionError there -- inside a prange loop, where what
// happens next is undefined: measured with Numba 0.65 / OpenMP layer on a point exactly on an edge with tolerance=0.0,
// the call raises for a batch of one, and for larger batches returns normally with the rows of the rest of the failing
// thread's chunk left zero (125 of 1000, 12 500 of 100 000).  The library takes the defined behaviour: the host turns
// the flag into ZeroDivisionError, always.
template <int MAXV>
CT_DEV void wachspress_weights(const Poly<MAXV> &polygon, P2 p, double tolerance, double (&w)[MAXV], int *zero_division = nullptr) {
    const int n = polygon.n;
    double w_sum = 0.0;
    P2 a = pget(polygon, n - 1);
    P2 b{polygon.x[0], polygon.y[0]};
    P2 U = to_vector(a, b);
    P2 V = to_vector(a, p);
    double Ai = fabs(cross_product(U, V));
    int ei = -1, ej = -1;  // on-edge case: linear interpolation between vertices ei, ej
    double ew = 0.0;
    if (within_perpendicular_distance(Ai, U, tolerance)) {
        ei = n - 1;
        ej = 0;
        P2 V2 = to_vector(a, p);
        const double length = sqrt(dot_product(U, U));
        if (length == 0.0 && zero_division) *zero_division = 1;
        ew = sqrt(dot_product(V2, V2)) / length;
    } else {
#pragma unroll
        for (int i = 0; i < MAXV; i++) {
            if (i >= n) break;
            int i_next = (i + 1 == n) ? 0 : i + 1;
            P2 c = pget(polygon, i_next);
            P2 W = to_vector(a, c);
            double Ci = fabs(cross_product(U, W));
            U = to_vector(b, c);
            V = to_vector(b, p);
            double Aj = fabs(cross_product(U, V));
            if (within_perpendicular_distance(Aj, U, tolerance)) {
                ei = i;
                ej = i_next;
                P2 V2 = to_vector(b, p);
                const double length = sqrt(dot_product(U, U));
                if (length == 0.0 && zero_division) *zero_division = 1;
                ew = sqrt(dot_product(V2, V2)) / length;
                break;
            }
            if ((Ai * Aj) == 0.0 && zero_division) *zero_division = 1;
            double wi = 2 * Ci / (Ai * Aj);
            w[i] = wi;
            w_sum += wi;
            a = b;
            b = c;
            Ai = Aj;
        }
    }
    if (ei >= 0) {
#pragma unroll
        for (int i = 0; i < MAXV; i++) w[i] = 0.0;
#pragma unroll
        for (int i = 0; i < MAXV; i++) {
            if (i == ei) w[i] = 1.0 - ew;
            if (i == ej) w[i] = ew;
        }
        return;
    }
    if (w_sum == 0.0 && zero_division) *zero_division = 1;
#pragma unroll
    for (int i = 0; i < MAXV; i++)
        if (i < n) w[i] /= w_sum;
}

}  // namespace ct
