/*
 * celltree_oracle.c -- CPU restatement of the numba_celltree query hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle for the CUDA
 * implementation in numba_celltree_b200/csrc.  Only tests/, the smoke check in
 * __graft_entry__.py and the cpu_baseline / --impl reference legs of bench.py
 * may load it; the product package never does.
 *
 * Parity status: PINNED.  Every entry point below is checked against outputs of
 * the reference itself (Deltares/numba_celltree v0.4.2 run under Numba in the
 * build container, fixtures in tests/golden/*.npz made by
 * tests/golden/make_golden.py) and against the known-answer tables of the
 * reference's own tests (re-stated in tests/test_oracle_*.py).
 *
 * Every function cites the reference file:line it follows (paths relative to
 * the reference repository root, package numba_celltree/).  Arithmetic is
 * written in the reference's source order; compile with -ffp-contract=off so
 * that no multiply-add is fused (Numba/LLVM emits none for this code).
 *
 * Numba lowers builtin min(a, b) to select(b < a, b, a) and max(a, b) to
 * select(b > a, b, a): ties and NaNs keep the FIRST argument.  nb_min / nb_max
 * below reproduce that; fmin/fmax would not.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define FILL_VALUE (-1)          /* constants.py:129 */
#define MAX_N_VERTEX 32          /* constants.py:128 */
#define INITIAL_STACK_LENGTH 32  /* constants.py:138 */
#define MIN_TOLERANCE 1e-15      /* constants.py:140 */
#define TOLERANCE_FACTOR 1e-12   /* constants.py:141 */
#define FLOAT_MAX DBL_MAX        /* constants.py:144 */
#define FLOAT_MIN (-DBL_MAX)     /* constants.py:143 (np.finfo.min) */

typedef struct { double x, y; } pt;                       /* Point / Vector, constants.py:36-43 */
typedef struct { double xmin, xmax, ymin, ymax; } box_t;  /* Box, constants.py:51-55 */

#pragma pack(push, 1)
typedef struct {                                          /* NodeDType, constants.py:91-106: 41 bytes */
    int64_t child;
    double Lmax;
    double Rmin;
    int64_t ptr;
    int64_t size;
    uint8_t dim;
} node_t;
#pragma pack(pop)

typedef struct {                                          /* Bucket, constants.py:73-79 */
    double Max, Min, Rmin, Lmax;
    int64_t index, size;
} bucket_t;

typedef struct {                                          /* CellTreeData, constants.py:82-89 */
    const int64_t *elements;
    int64_t n_elem;
    int64_t n_max_vert;
    const double *vertices;
    const node_t *nodes;
    const int64_t *bb_indices;
    const double *bb_coords;
    double bbox[4];
} tree_t;

static inline double nb_min(double a, double b) { return b < a ? b : a; }
static inline double nb_max(double a, double b) { return b > a ? b : a; }

/* ------------------------------------------------------------------ */
/* growable int stack: utils.py:19-44                                  */
typedef struct { int64_t *data; int64_t cap; int64_t size; int64_t inline_buf[INITIAL_STACK_LENGTH]; } istack;
static inline void stack_init(istack *s) { s->data = s->inline_buf; s->cap = INITIAL_STACK_LENGTH; s->size = 0; }
static inline void stack_free(istack *s) { if (s->data != s->inline_buf) free(s->data); }
static inline void stack_push(istack *s, int64_t v)       /* utils.py:35-39, grow :90-97 */
{
    if (s->size >= s->cap) {
        int64_t *nd = (int64_t *)malloc(sizeof(int64_t) * 2 * s->cap);
        memcpy(nd, s->data, sizeof(int64_t) * s->cap);
        if (s->data != s->inline_buf) free(s->data);
        s->data = nd;
        s->cap *= 2;
    }
    s->data[s->size++] = v;
}
static inline int64_t stack_pop(istack *s) { return s->data[--s->size]; }  /* utils.py:30-32 */

/* ------------------------------------------------------------------ */
/* geometry_utils.py scalar helpers                                    */
static inline pt to_vector(pt a, pt b) { pt v = { b.x - a.x, b.y - a.y }; return v; }    /* :23-25 */
static inline double cross_product(pt u, pt v) { return u.x * v.y - u.y * v.x; }         /* :57-59 */
static inline double dot_product(pt u, pt v) { return u.x * v.x + u.y * v.y; }           /* :62-64 */
static inline double length_squared(pt v) { return v.x * v.x + v.y * v.y; }              /* :67-69 */
static inline pt to_point(double t, pt a, pt V) { pt p = { a.x + t * V.x, a.y + t * V.y }; return p; } /* :52-54 */

static inline int polygon_length(const int64_t *face, int n)                              /* :72-79 */
{
    for (int i = 3; i < n; i++)
        if (face[i] == FILL_VALUE) return i;
    return n;
}

static double polygon_area(const pt *polygon, int length)                                 /* :82-95 */
{
    double area = 0.0;
    pt a = polygon[0];
    pt b = polygon[1];
    pt U = to_vector(a, b);
    for (int i = 2; i < length; i++) {
        pt c = polygon[i];
        pt V = to_vector(c, a);
        area += fabs(cross_product(U, V));
        b = c;
        U = V;
    }
    (void)b;
    return 0.5 * area;
}

static inline int within_perpendicular_distance(double UxV, pt U, double tolerance)       /* :150-166 */
{
    return (UxV * UxV) < ((tolerance * length_squared(U)) * tolerance);
}

static inline int in_bounds(pt p, pt a, pt b)                                             /* :169-192 */
{
    double xmin = nb_min(a.x, b.x);
    double xmax = nb_max(a.x, b.x);
    double ymin = nb_min(a.y, b.y);
    double ymax = nb_max(a.y, b.y);
    double dx = xmax - xmin;
    double dy = ymax - ymin;
    int use_x_bound = fabs(dx) >= fabs(dy);
    return (use_x_bound && ((p.x >= xmin) && (p.x <= xmax))) ||
           (!use_x_bound && ((p.y >= ymin) && (p.y <= ymax)));
}

static int point_in_polygon_or_on_edge(pt p, const pt *poly, int length, double tolerance) /* :195-223 */
{
    pt v0 = poly[length - 1];
    pt U = to_vector(p, v0);
    int c = 0;
    for (int i = 0; i < length; i++) {
        pt v1 = poly[i];
        if (v1.x == v0.x && v1.y == v0.y) continue;
        pt V = to_vector(p, v1);
        double A = cross_product(U, V);
        pt W = to_vector(v0, v1);
        if (within_perpendicular_distance(A, W, tolerance) && in_bounds(p, v0, v1)) return 1;
        if (((v0.y > p.y) != (v1.y > p.y)) &&
            (p.x < ((v1.x - v0.x) * (p.y - v0.y) / (v1.y - v0.y) + v0.x)))
            c = !c;
        v0 = v1;
        U = V;
    }
    return c;
}

static int point_on_edge(pt p, pt v0, pt v1, double tolerance)                            /* :226-238 */
{
    if (v1.x == v0.x && v1.y == v0.y) return 0;
    pt U = to_vector(p, v0);
    pt V = to_vector(p, v1);
    pt W = to_vector(v0, v1);
    double A = cross_product(U, V);
    if (within_perpendicular_distance(A, W, tolerance) && in_bounds(p, v0, v1)) return 1;
    return 0;
}

static inline int boxes_intersect(box_t a, box_t b)                                       /* :292-300 */
{
    return a.xmin < b.xmax && b.xmin < a.xmax && a.ymin < b.ymax && b.ymin < a.ymax;
}

static inline int left_of(pt a, pt p, pt U) { return U.x * (a.y - p.y) > U.y * (a.x - p.x); } /* :318-322 */

static inline int has_overlap(double a, double b, double p, double q, double tolerance)   /* :325-329 */
{
    return ((nb_min(a, b) - nb_max(p, q)) < tolerance) && ((nb_min(p, q) - nb_max(a, b)) < tolerance);
}

static void intersection_location_point(pt V, pt U, pt a, pt p, double tolerance, double *x, double *y) /* :332-346 */
{
    double denom = cross_product(V, U);
    if (within_perpendicular_distance(denom, V, tolerance)) { *x = NAN; *y = NAN; return; }
    pt R = to_vector(a, p);
    double t = cross_product(R, U) / denom;
    *x = a.x + t * V.x;
    *y = a.y + t * V.y;
}

static void midpoint_collinear_lines(pt a, pt b, pt p, pt q, double *x, double *y)        /* :349-374 */
{
    if (a.x > b.x || (a.x == b.x && a.y > b.y)) { pt t = a; a = b; b = t; }
    if (p.x > q.x || (p.x == q.x && p.y > q.y)) { pt t = p; p = q; q = t; }
    double overlap_start_x = nb_max(a.x, p.x);
    double overlap_start_y = nb_max(a.y, p.y);
    double overlap_end_x = nb_min(b.x, q.x);
    double overlap_end_y = nb_min(b.y, q.y);
    if (overlap_start_x > overlap_end_x || overlap_start_y > overlap_end_y) { *x = NAN; *y = NAN; return; }
    *x = 0.5 * (overlap_start_x + overlap_end_x);
    *y = 0.5 * (overlap_start_y + overlap_end_y);
}

static int lines_intersect(pt a, pt b, pt p, pt q, double *x, double *y)                  /* :377-418 */
{
    pt V = to_vector(a, b);
    pt U = to_vector(p, q);
    *x = NAN; *y = NAN;
    if ((U.x == 0 && U.y == 0) || (V.x == 0 && V.y == 0)) return 0;
    if ((U.x == 0) && (V.x == 0) && a.x != p.x) return 0;
    if ((U.y == 0) && (V.y == 0) && a.y != p.y) return 0;
    double tolerance = nb_max(MIN_TOLERANCE, TOLERANCE_FACTOR * nb_max(fabs(U.x), fabs(U.y)));
    if ((!has_overlap(a.x, b.x, p.x, q.x, tolerance)) || (!has_overlap(a.y, b.y, p.y, q.y, tolerance))) return 0;
    if ((left_of(a, p, U) != left_of(b, p, U)) && (left_of(p, a, V) != left_of(q, a, V))) {
        intersection_location_point(V, U, a, p, tolerance, x, y);
        return 1;
    }
    pt R = to_vector(a, p);
    pt S = to_vector(a, q);
    if (within_perpendicular_distance(cross_product(V, R), V, tolerance) &&
        within_perpendicular_distance(cross_product(V, S), V, tolerance)) {
        midpoint_collinear_lines(a, b, p, q, x, y);
        return 1;
    }
    return 0;
}

/* copy_vertices / copy_vertices_into: geometry_utils.py:490-510 */
static inline int copy_vertices(const double *vertices, const int64_t *face, int n_max_vert, pt *out)
{
    int length = polygon_length(face, n_max_vert);
    for (int i = 0; i < length; i++) {
        const double *v = vertices + 2 * face[i];
        out[i].x = v[0];
        out[i].y = v[1];
    }
    return length;
}

static inline void copy_box_vertices(box_t box, pt *a)                                    /* :513-524 */
{
    a[0].x = box.xmin; a[0].y = box.ymin;
    a[1].x = box.xmax; a[1].y = box.ymin;
    a[2].x = box.xmax; a[2].y = box.ymax;
    a[3].x = box.xmin; a[3].y = box.ymax;
}

static inline box_t as_box(const double *a) { box_t b = { a[0], a[1], a[2], a[3] }; return b; }  /* :33-40 */

/* ------------------------------------------------------------------ */
/* algorithms/cohen_sutherland.py                                      */
enum { CS_INSIDE = 0, CS_LEFT = 1, CS_RIGHT = 2, CS_LOWER = 4, CS_UPPER = 8 };

static inline int get_clip(pt a, box_t box)                                               /* cohen_sutherland.py:18-33 */
{
    int p = CS_INSIDE;
    if (a.x < box.xmin) p |= CS_LEFT;
    else if (a.x > box.xmax) p |= CS_RIGHT;
    if (a.y < box.ymin) p |= CS_LOWER;
    else if (a.y > box.ymax) p |= CS_UPPER;
    return p;
}

/* returns 1/0; -1 on "Undefined clipping state" (cohen_sutherland.py:86) */
static int cohen_sutherland_line_box_clip(pt a, pt b, box_t box, pt *c, pt *d)            /* :36-101 */
{
    pt nanp = { NAN, NAN };
    *c = nanp; *d = nanp;
    double dx = b.x - a.x;
    double dy = b.y - a.y;
    if (dx == 0.0 && dy == 0.0) return 0;
    int k1 = get_clip(a, box);
    int k2 = get_clip(b, box);
    while ((k1 | k2) != CS_INSIDE) {
        if ((k1 & k2) != 0) return 0;
        int opt = k1 ? k1 : k2;
        double x, y;
        if (opt & CS_UPPER) { x = a.x + dx * (box.ymax - a.y) / dy; y = box.ymax; }
        else if (opt & CS_LOWER) { x = a.x + dx * (box.ymin - a.y) / dy; y = box.ymin; }
        else if (opt & CS_RIGHT) { y = a.y + dy * (box.xmax - a.x) / dx; x = box.xmax; }
        else if (opt & CS_LEFT) { y = a.y + dy * (box.xmin - a.x) / dx; x = box.xmin; }
        else return -1;
        if (opt == k1) { a.x = x; a.y = y; k1 = get_clip(a, box); }
        else if (opt == k2) { b.x = x; b.y = y; k2 = get_clip(b, box); }
        dx = b.x - a.x;
        dy = b.y - a.y;
        if (dx == 0.0 && dy == 0.0) return 0;
    }
    *c = a; *d = b;
    return 1;
}

/* ------------------------------------------------------------------ */
/* algorithms/cyrus_beck.py                                            */
static inline int cb_compute_intersection(pt a, pt s, pt v0, pt v1, double *t)            /* cyrus_beck.py:35-52 */
{
    pt si = to_vector(a, v0);
    pt n = { -(v1.y - v0.y), (v1.x - v0.x) };
    double n_si = dot_product(n, si);
    double k = dot_product(n, s);
    *t = n_si / k;
    return n_si > 0;
}

static inline void cb_intersections(pt a, pt s, const pt *poly, int length, int i0, int i1, double *t0o, double *t1o) /* :55-72 */
{
    /* i1 may be -1 (single crossing): Python negative indexing => last vertex; (i1+1)%length == 0 */
    pt v0 = poly[i0];
    pt v01 = poly[(i0 + 1) % length];
    pt v1 = poly[i1 < 0 ? i1 + length : i1];
    pt v11 = poly[(i1 + 1) % length];
    double t0, t1;
    (void)cb_compute_intersection(a, s, v0, v01, &t0);
    int enters1 = cb_compute_intersection(a, s, v1, v11, &t1);
    if (enters1) { *t0o = t1; *t1o = t0; }
    else { *t0o = t0; *t1o = t1; }
}

static inline int cb_overlap(double ta, double tb, double t0, double t1)                  /* :75-82 */
{
    if (ta > tb) { double t = ta; ta = tb; tb = t; }
    if (t0 > t1) { double t = t0; t0 = t1; t1 = t; }
    double vector_overlap = nb_max(0.0, nb_min(tb, t1) - nb_max(ta, t0));
    return vector_overlap > 0.0;
}

static inline int cb_aligned(pt U, pt V)                                                  /* :85-100 */
{
    if ((U.x == 0 && U.y == 0) || (V.x == 0 && V.y == 0)) return 1;
    if (U.x != 0 && V.x != 0) return (U.x > 0) == (V.x > 0);
    if (U.y != 0 && V.y != 0) return (U.y > 0) == (V.y > 0);
    return 0;
}

static int cb_collinear_case(pt a, pt b, pt v0, pt v1, pt *c, pt *d)                      /* :103-139 */
{
    pt nanp = { NAN, NAN };
    pt _b = { b.x - a.x, b.y - a.y };
    pt _v0 = { v0.x - a.x, v0.y - a.y };
    pt _v1 = { v1.x - a.x, v1.y - a.y };
    pt U = _b;
    pt V = to_vector(_v0, _v1);
    if (!cb_aligned(U, V)) {
        pt t = v0; v0 = v1; v1 = t;
        t = _v0; _v0 = _v1; _v1 = t;
    }
    pt n = { -_b.y, _b.x };
    double ta = 0.0;
    double tb = cross_product(n, _b);
    double t0 = cross_product(n, _v0);
    double t1 = cross_product(n, _v1);
    if (!cb_overlap(ta, tb, t0, t1)) { *c = nanp; *d = nanp; return 0; }
    *c = (t0 < ta) ? v0 : a;
    *d = (t1 > tb) ? v1 : b;
    return 1;
}

static int cyrus_beck_line_polygon_clip(pt a, pt b, const pt *poly, int length, double tolerance, pt *c, pt *d) /* :143-241 */
{
    pt nanp = { NAN, NAN };
    *c = nanp; *d = nanp;
    pt s = to_vector(a, b);
    if (s.x == 0 && s.y == 0) return 0;
    int a_inside = point_in_polygon_or_on_edge(a, poly, length, tolerance);
    int b_inside = point_in_polygon_or_on_edge(b, poly, length, tolerance);
    if (a_inside && b_inside) { *c = a; *d = b; return 1; }

    int i0 = -1, i1 = -1, i = 0, k = 0;
    pt v = poly[0];
    double ksi = cross_product(to_vector(a, v), s);
    while (i < length && k < 2) {
        pt v0 = poly[i];
        pt v1 = poly[(i + 1) % length];
        double eta = cross_product(to_vector(a, v1), s);
        if ((ksi < 0.0) ^ (eta < 0.0)) {
            if (k == 0) i0 = i; else i1 = i;
            k += 1;
        } else if ((ksi == 0.0) && (eta == 0.0)) {
            return cb_collinear_case(a, b, v0, v1, c, d);
        }
        ksi = eta;
        i += 1;
    }
    if (k == 0) return 0;

    double t0, t1;
    cb_intersections(a, s, poly, length, i0, i1, &t0, &t1);
    if (t0 == t1) {
        if (a_inside && t1 != 0.0) t0 = 0.0;
        else if (b_inside && t0 != 1.0) t1 = 1.0;
        else return 0;
    }
    if (t1 < t0) { double t = t0; t0 = t1; t1 = t; }
    int valid0 = t0 >= 0 && t0 < 1;
    int valid1 = t1 > 0 && t1 <= 1;
    if (valid0 && valid1) { *c = to_point(t0, a, s); *d = to_point(t1, a, s); return 1; }
    else if (valid0) { *c = to_point(t0, a, s); *d = b; return 1; }
    else if (valid1) { *c = a; *d = to_point(t1, a, s); return 1; }
    return 0;
}

/* ------------------------------------------------------------------ */
/* algorithms/sutherland_hodgman.py                                    */
static inline int sh_inside(pt p, pt r, pt U) { return U.x * (p.y - r.y) > U.y * (p.x - r.x); }  /* :56-60 */

static inline int sh_intersection(pt a, pt V, pt r, pt N, pt *out)                        /* :63-74 */
{
    pt W = { r.x - a.x, r.y - a.y };
    double nw = dot_product(N, W);
    double nv = dot_product(N, V);
    if (nv != 0) {
        double t = nw / nv;
        out->x = a.x + t * V.x;
        out->y = a.y + t * V.y;
        return 1;
    }
    out->x = NAN; out->y = NAN;
    return 0;
}

static double polygon_polygon_clip_area(const pt *polygon, int n_polygon, const pt *clipper, int n_clip) /* :84-148 */
{
    pt subject[2 * MAX_N_VERTEX];
    pt output[2 * MAX_N_VERTEX];
    int n_output = n_polygon;
    for (int i = 0; i < n_output; i++) output[i] = polygon[i];

    pt r = clipper[n_clip - 1];
    for (int i = 0; i < n_clip; i++) {
        pt s = clipper[i];
        pt U = { s.x - r.x, s.y - r.y };
        if (U.x == 0 && U.y == 0) continue;
        pt N = { -U.y, U.x };
        int length = n_output;
        for (int j = 0; j < length; j++) subject[j] = output[j];
        n_output = 0;
        pt a = subject[length - 1];
        int a_inside = sh_inside(a, r, U);
        for (int j = 0; j < length; j++) {
            pt b = subject[j];
            pt V = { b.x - a.x, b.y - a.y };
            if (V.x == 0 && V.y == 0) continue;
            int b_inside = sh_inside(b, r, U);
            if (b_inside) {
                if (!a_inside) {
                    pt point;
                    if (sh_intersection(a, V, r, N, &point)) output[n_output++] = point;
                }
                output[n_output++] = b;
            } else if (a_inside) {
                pt point;
                if (sh_intersection(a, V, r, N, &point)) output[n_output++] = point;
                else { b_inside = 1; output[n_output++] = b; }
            }
            a = b;
            a_inside = b_inside;
        }
        if (n_output < 3) return 0.0;
        r = s;
    }
    return polygon_area(output, n_output);
}

/* algorithms/separating_axis.py */
static inline void extrema_projected(pt norm, const pt *polygon, int length, double *mn, double *mx) /* :17-27 */
{
    double min_proj = FLOAT_MAX, max_proj = FLOAT_MIN;
    for (int i = 0; i < length; i++) {
        double proj = dot_product(polygon[i], norm);
        min_proj = nb_min(min_proj, proj);
        max_proj = nb_max(max_proj, proj);
    }
    *mn = min_proj; *mx = max_proj;
}

static inline int is_separating_axis(pt norm, const pt *a, const pt *b, int la, int lb)   /* :30-39 */
{
    double mina, maxa, minb, maxb;
    extrema_projected(norm, a, la, &mina, &maxa);
    extrema_projected(norm, b, lb, &minb, &maxb);
    if (maxa > minb && maxb > mina) return 0;
    return 1;
}

static int separating_axes(const pt *a, int la, const pt *b, int lb)                      /* :42-55 */
{
    pt p = a[la - 1];
    for (int i = 0; i < la; i++) {
        pt q = a[i];
        pt norm = { p.y - q.y, q.x - p.x };
        p = q;
        if (norm.x == 0.0 && norm.y == 0.0) continue;
        if (is_separating_axis(norm, a, b, la, lb)) return 0;
    }
    return 1;
}

/* algorithms/barycentric_triangle.py:28-43 */
static inline void tri_compute_weights(pt a, pt b, pt c, pt p, double *weights)
{
    pt ab = to_vector(a, b);
    pt ac = to_vector(a, c);
    pt ap = to_vector(a, p);
    double Aa = fabs(cross_product(ab, ap));
    double Ac = fabs(cross_product(ac, ap));
    double A = fabs(cross_product(ab, ac));
    double inv_denom = 1.0 / A;
    double w = inv_denom * Aa;
    double v = inv_denom * Ac;
    double u = 1.0 - v - w;
    weights[0] = u;
    weights[1] = v;
    weights[2] = w;
}

/* algorithms/barycentric_wachspress.py:26-35 */
static inline void interp_edge_case(pt a, pt U, pt p, double *weights, int n_w, int i, int j)
{
    for (int k = 0; k < n_w; k++) weights[k] = 0;
    pt V = to_vector(a, p);
    double w = sqrt(dot_product(V, V)) / sqrt(dot_product(U, U));
    weights[i] = 1.0 - w;
    weights[j] = w;
}

/* algorithms/barycentric_wachspress.py:38-85 */
static void wachspress_compute_weights(const pt *polygon, int n, pt p, double *weights, int n_w, double tolerance)
{
    double w_sum = 0.0;
    pt a = polygon[n - 1];
    pt b = polygon[0];
    pt U = to_vector(a, b);
    pt V = to_vector(a, p);
    double Ai = fabs(cross_product(U, V));
    if (within_perpendicular_distance(Ai, U, tolerance)) {
        interp_edge_case(a, U, p, weights, n_w, n - 1, 0);
        return;
    }
    for (int i = 0; i < n; i++) {
        int i_next = (i + 1) % n;
        pt c = polygon[i_next];
        pt W = to_vector(a, c);
        double Ci = fabs(cross_product(U, W));
        U = to_vector(b, c);
        V = to_vector(b, p);
        double Aj = fabs(cross_product(U, V));
        if (within_perpendicular_distance(Aj, U, tolerance)) {
            interp_edge_case(b, U, p, weights, n_w, i, i_next);
            return;
        }
        double w = 2 * Ci / (Ai * Aj);
        weights[i] = w;
        w_sum += w;
        a = b;
        b = c;
        Ai = Aj;
    }
    for (int i = 0; i < n; i++) weights[i] /= w_sum;
}

/* ================================================================== */
/* Exported scalar entry points (known-answer tests drive these)       */
/* ================================================================== */
int orc_point_in_polygon_or_on_edge(double px, double py, const double *poly, int length, double tol)
{
    pt p = { px, py };
    return point_in_polygon_or_on_edge(p, (const pt *)poly, length, tol);
}

int orc_cohen_sutherland(const double *ab, const double *box, double *cd)
{
    pt a = { ab[0], ab[1] }, b = { ab[2], ab[3] }, c, d;
    int r = cohen_sutherland_line_box_clip(a, b, as_box(box), &c, &d);
    cd[0] = c.x; cd[1] = c.y; cd[2] = d.x; cd[3] = d.y;
    return r;
}

int orc_cyrus_beck(const double *ab, const double *poly, int length, double tol, double *cd)
{
    pt a = { ab[0], ab[1] }, b = { ab[2], ab[3] }, c, d;
    int r = cyrus_beck_line_polygon_clip(a, b, (const pt *)poly, length, tol, &c, &d);
    cd[0] = c.x; cd[1] = c.y; cd[2] = d.x; cd[3] = d.y;
    return r;
}

int orc_lines_intersect(const double *ab, const double *pq, double *xy)
{
    pt a = { ab[0], ab[1] }, b = { ab[2], ab[3] }, p = { pq[0], pq[1] }, q = { pq[2], pq[3] };
    return lines_intersect(a, b, p, q, &xy[0], &xy[1]);
}

double orc_clip_area(const double *polygon, int n_polygon, const double *clipper, int n_clip)
{
    return polygon_polygon_clip_area((const pt *)polygon, n_polygon, (const pt *)clipper, n_clip);
}

int orc_separating_axes(const double *a, int la, const double *b, int lb)
{
    return separating_axes((const pt *)a, la, (const pt *)b, lb);
}

void orc_wachspress(const double *polygon, int n, double px, double py, double *weights, int n_w, double tol)
{
    pt p = { px, py };
    wachspress_compute_weights((const pt *)polygon, n, p, weights, n_w, tol);
}

/* ------------------------------------------------------------------ */
/* Exported helpers off the API path, batched        */
/* ------------------------------------------------------------------ */
static inline int point_inside_box(pt a, box_t box)                                      /* geometry_utils.py:527-529 */
{
    return box.xmin < a.x && a.x < box.xmax && box.ymin < a.y && a.y < box.ymax;
}

/* algorithms/liang_barsky.py:10-64: parametric clip of a -> b against the four box sides, taken in the order
 * left, right, lower, upper; a segment of zero length, or one that only touches (t0 == t1), is a miss. */
static int liang_barsky_line_box_clip(pt a, pt b, box_t box, pt *c, pt *d)
{
    const double nan_ = NAN;
    c->x = c->y = d->x = d->y = nan_;
    double dx = b.x - a.x, dy = b.y - a.y;
    if (dx == 0.0 && dy == 0.0) return 0;
    if (point_inside_box(a, box) && point_inside_box(b, box)) { *c = a; *d = b; return 1; }
    double t0 = 0.0, t1 = 1.0;
    const double P[4] = { -dx, dx, -dy, dy };
    const double Q[4] = { a.x - box.xmin, box.xmax - a.x, a.y - box.ymin, box.ymax - a.y };
    for (int i = 0; i < 4; i++) {
        double p_i = P[i], q_i = Q[i];
        if (p_i == 0) {
            if (q_i < 0) return 0;              /* parallel to this side and outside of it */
        } else {
            double t = q_i / p_i;
            if (p_i < 0) {
                if (t > t1) return 0;
                else if (t > t0) t0 = t;
            } else if (p_i > 0) {
                if (t < t0) return 0;
                else if (t < t1) t1 = t;
            }
        }
    }
    if (t0 == t1) return 0;
    c->x = a.x + t0 * dx; c->y = a.y + t0 * dy;
    d->x = a.x + t1 * dx; d->y = a.y + t1 * dy;
    return 1;
}

void orc_liang_barsky(const double *a, const double *b, const double *boxes, int64_t n, uint8_t *hit, double *c, double *d)
{
    for (int64_t i = 0; i < n; i++) {
        pt pa = { a[2 * i], a[2 * i + 1] }, pb = { b[2 * i], b[2 * i + 1] }, pc, pd;
        hit[i] = (uint8_t)liang_barsky_line_box_clip(pa, pb, as_box(boxes + 4 * i), &pc, &pd);
        c[2 * i] = pc.x; c[2 * i + 1] = pc.y; d[2 * i] = pd.x; d[2 * i + 1] = pd.y;
    }
}

void orc_cohen_sutherland_batch(const double *a, const double *b, const double *boxes, int64_t n, uint8_t *hit, double *c, double *d)
{
    for (int64_t i = 0; i < n; i++) {
        pt pa = { a[2 * i], a[2 * i + 1] }, pb = { b[2 * i], b[2 * i + 1] }, pc, pd;
        hit[i] = (uint8_t)cohen_sutherland_line_box_clip(pa, pb, as_box(boxes + 4 * i), &pc, &pd);
        c[2 * i] = pc.x; c[2 * i + 1] = pc.y; d[2 * i] = pd.x; d[2 * i + 1] = pd.y;
    }
}

void orc_cyrus_beck_batch(const double *a, const double *b, int64_t n, const double *poly, int length, double tol,
                          uint8_t *hit, double *c, double *d)
{
    for (int64_t i = 0; i < n; i++) {
        pt pa = { a[2 * i], a[2 * i + 1] }, pb = { b[2 * i], b[2 * i + 1] }, pc, pd;
        hit[i] = (uint8_t)cyrus_beck_line_polygon_clip(pa, pb, (const pt *)poly, length, tol, &pc, &pd);
        c[2 * i] = pc.x; c[2 * i + 1] = pc.y; d[2 * i] = pd.x; d[2 * i + 1] = pd.y;
    }
}

/* geometry_utils.py:98-147: the plain crossing-number test (no tolerance, no on-edge acceptance) */
static int point_in_polygon(pt p, const pt *poly, int length)
{
    pt v0 = poly[length - 1];
    int inside = 0;
    for (int i = 0; i < length; i++) {
        pt v1 = poly[i];
        if (((v0.y > p.y) != (v1.y > p.y)) && (p.x < ((v1.x - v0.x) * (p.y - v0.y) / (v1.y - v0.y) + v0.x))) inside = !inside;
        v0 = v1;
    }
    return inside;
}

void orc_points_in_polygon(const double *points, int64_t n, const double *poly, int length, uint8_t *inside)
{
    for (int64_t i = 0; i < n; i++) {
        pt p = { points[2 * i], points[2 * i + 1] };
        inside[i] = (uint8_t)point_in_polygon(p, (const pt *)poly, length);
    }
}

/* geometry_utils.py:241-270: half-plane signs first, then the three on-edge tests */
static int point_in_triangle(pt p, pt ta, pt tb, pt tc, double tolerance)
{
    pt ap = to_vector(ta, p), bp = to_vector(tb, p), cp = to_vector(tc, p);
    pt ab = to_vector(ta, tb), bc = to_vector(tb, tc), ca = to_vector(tc, ta);
    double A = cross_product(ab, ap), B = cross_product(bc, bp), C = cross_product(ca, cp);
    int sA = A > 0, sB = B > 0, sC = C > 0;
    if (sA == sB && sB == sC) return 1;
    if ((within_perpendicular_distance(A, ab, tolerance) && in_bounds(p, ta, tb))
        || (within_perpendicular_distance(B, bc, tolerance) && in_bounds(p, tb, tc))
        || (within_perpendicular_distance(C, ca, tolerance) && in_bounds(p, tc, ta)))
        return 1;
    return 0;
}

/* geometry_utils.py:273-289 */
void orc_points_in_triangles(const double *points, const int64_t *face_indices, int64_t n, const int64_t *faces,
                             int64_t n_max_vert, const double *vertices, double tolerance, uint8_t *inside)
{
    const pt *v = (const pt *)vertices;
    for (int64_t i = 0; i < n; i++) {
        const int64_t *face = faces + face_indices[i] * n_max_vert;
        pt p = { points[2 * i], points[2 * i + 1] };
        inside[i] = (uint8_t)point_in_triangle(p, v[face[0]], v[face[1]], v[face[2]], tolerance);
    }
}

/* ================================================================== */
/* Mesh preparation: geometry_utils.py                                 */
/* ================================================================== */
static inline void flip(int64_t *face, int length)                                        /* :532-538 */
{
    int end = length - 1;
    for (int i = 0; i < (int)(length / 2.0); i++) {
        int j = end - i;
        int64_t t = face[i]; face[i] = face[j]; face[j] = t;
    }
}

/* geometry_utils.py:541-561.  Literal: after flip() the loop carries on with the
 * stale a, b (the reference does too). */
void orc_counter_clockwise(const double *vertices, int64_t *faces, int64_t n_face, int n_max_vert)
{
#pragma omp parallel for schedule(static)
    for (int64_t i_face = 0; i_face < n_face; i_face++) {
        int64_t *face = faces + i_face * n_max_vert;
        int length = polygon_length(face, n_max_vert);
        pt a = { vertices[2 * face[length - 2]], vertices[2 * face[length - 2] + 1] };
        pt b = { vertices[2 * face[length - 1]], vertices[2 * face[length - 1] + 1] };
        for (int i = 0; i < length; i++) {
            pt c = { vertices[2 * face[i]], vertices[2 * face[i] + 1] };
            pt u = to_vector(a, b);
            pt v = to_vector(a, c);
            double product = cross_product(u, v);
            if (product == 0) { a = b; b = c; }
            else if (product < 0) flip(face, length);
            else break;
        }
    }
}

/* geometry_utils.py:421-456 (serial in the reference: no parallel=True) */
void orc_build_face_bboxes(const int64_t *faces, const double *vertices, int64_t n_face, int n_max_vert, double *bbox_coords)
{
    for (int64_t i = 0; i < n_face; i++) {
        const int64_t *polygon = faces + i * n_max_vert;
        const double *first = vertices + 2 * polygon[0];
        double xmin = first[0], xmax = first[0], ymin = first[1], ymax = first[1];
        for (int k = 1; k < n_max_vert; k++) {
            int64_t index = polygon[k];
            if (index == FILL_VALUE) break;
            double x = vertices[2 * index], y = vertices[2 * index + 1];
            xmin = nb_min(xmin, x);
            xmax = nb_max(xmax, x);
            ymin = nb_min(ymin, y);
            ymax = nb_max(ymax, y);
        }
        bbox_coords[4 * i + 0] = xmin;
        bbox_coords[4 * i + 1] = xmax;
        bbox_coords[4 * i + 2] = ymin;
        bbox_coords[4 * i + 3] = ymax;
    }
}

/* geometry_utils.py:459-487 */
void orc_build_edge_bboxes(const int64_t *edges, const double *vertices, int64_t n_edge, double tolerance, double *bbox_coords)
{
    for (int64_t i = 0; i < n_edge; i++) {
        double x0 = vertices[2 * edges[2 * i]], y0 = vertices[2 * edges[2 * i] + 1];
        double x1 = vertices[2 * edges[2 * i + 1]], y1 = vertices[2 * edges[2 * i + 1] + 1];
        bbox_coords[4 * i + 0] = nb_min(x0 - tolerance, x1 - tolerance);
        bbox_coords[4 * i + 1] = nb_max(x0 + tolerance, x1 + tolerance);
        bbox_coords[4 * i + 2] = nb_min(y0 - tolerance, y1 - tolerance);
        bbox_coords[4 * i + 3] = nb_max(y0 + tolerance, y1 + tolerance);
    }
}

/* ================================================================== */
/* Tree construction: creation.py                                      */
/* ================================================================== */
int64_t orc_pessimistic_n_nodes(int64_t n_elements)                                       /* creation.py:216-230 */
{
    int64_t n_nodes = n_elements;
    int64_t nodes = (int64_t)ceil(n_elements / 2.0);
    while (nodes > 1) {
        n_nodes += nodes;
        nodes = (int64_t)ceil(nodes / 2.0);
    }
    return n_nodes + 1;
}

static inline int centroid_test(const bucket_t *bucket, const double *box, int dim)      /* creation.py:44-51 */
{
    double centroid = box[2 * dim] + 0.5 * (box[2 * dim + 1] - box[2 * dim]);
    return (centroid >= bucket->Min) && (centroid < bucket->Max);
}

static int64_t stable_partition(int64_t *bb_indices, const double *bb_coords, int64_t begin, int64_t end,
                                const bucket_t *bucket, int dim, int64_t *temp)          /* creation.py:54-112 */
{
    int64_t n = end - begin;
    int64_t count_true = 0, count_false = 0;  /* false group is filled from the back of temp */
    for (int64_t k = begin; k < end; k++) {
        int64_t i = bb_indices[k];
        if (centroid_test(bucket, bb_coords + 4 * i, dim)) temp[count_true++] = i;
        else { temp[n - 1 - count_false] = i; count_false++; }
    }
    for (int64_t i = 0; i < count_true; i++) bb_indices[begin + i] = temp[i];
    int64_t start_second = begin + count_true;
    for (int64_t i = 0; i < count_false; i++) bb_indices[start_second + i] = temp[n - 1 - i];
    return start_second;
}

/* creation.py:115-150.  Returns 0, or -1 when an element's centroid falls in no
 * bucket (the reference indexes past its bucket list there: IndexError). */
static int sort_bbox_indices(int64_t *bb_indices, const double *bb_coords, bucket_t *buckets, int n_buckets,
                             int64_t ptr, int64_t size, int dim, int64_t *temp)
{
    int64_t current = ptr;
    int64_t end = ptr + size;
    buckets[0].index = ptr;
    int i = 1;
    while (current != end) {
        if (i - 1 >= n_buckets) return -1;
        bucket_t *bucket = &buckets[i - 1];
        current = stable_partition(bb_indices, bb_coords, current, end, bucket, dim, temp);
        int64_t start = bucket->index;
        bucket->size = current - start;
        if (i < n_buckets) buckets[i].index = buckets[i - 1].index + buckets[i - 1].size;
        i += 1;
    }
    return 0;
}

static void get_bounds(int64_t index, int64_t size, const double *bb_coords, const int64_t *bb_indices, int dim,
                       double *Rmin_o, double *Lmax_o)                                   /* creation.py:153-171 */
{
    double Rmin = FLOAT_MAX, Lmax = FLOAT_MIN;
    for (int64_t i = index; i < index + size; i++) {
        int64_t data_index = bb_indices[i];
        double value = bb_coords[4 * data_index + 2 * dim];
        if (value < Rmin) Rmin = value;
        value = bb_coords[4 * data_index + 2 * dim + 1];
        if (value > Lmax) Lmax = value;
    }
    *Rmin_o = Rmin; *Lmax_o = Lmax;
}

static void split_plane(const bucket_t *buckets, int n, int64_t root_size, double range_Lmax, double range_Rmin,
                        double bucket_length, int *plane_o, double *Lmax_o, double *Rmin_o) /* creation.py:174-213 */
{
    double plane_min_cost = FLOAT_MAX;
    int plane = INT32_MAX;
    int64_t bbs_in_left = 0, bbs_in_right = 0;
    for (int i = 1; i < n; i++) {
        const bucket_t *current_bucket = &buckets[i - 1];
        const bucket_t *next_bucket = &buckets[i];
        bbs_in_left += current_bucket->size;
        bbs_in_right = root_size - bbs_in_left;
        double left_volume = (current_bucket->Lmax - range_Rmin) / bucket_length;
        double right_volume = (range_Lmax - next_bucket->Rmin) / bucket_length;
        double plane_cost = left_volume * (double)bbs_in_left + right_volume * (double)bbs_in_right;
        if (plane_cost < plane_min_cost) { plane_min_cost = plane_cost; plane = i; }
    }
    double Lmax = FLOAT_MIN, Rmin = FLOAT_MAX;
    for (int i = 0; i < plane && i < n; i++) if (buckets[i].Lmax > Lmax) Lmax = buckets[i].Lmax;
    for (int i = plane; i < n; i++) if (buckets[i].Rmin < Rmin) Rmin = buckets[i].Rmin;
    *plane_o = plane; *Lmax_o = Lmax; *Rmin_o = Rmin;
}

static inline void set_node(node_t *nd, int64_t ptr, int64_t size, int dim)               /* creation.py:27-41 */
{
    nd->child = -1; nd->Lmax = -1.0; nd->Rmin = -1.0; nd->ptr = ptr; nd->size = size; nd->dim = (uint8_t)(dim ? 1 : 0);
}

/* creation.py:233-381 (build) + :384-413 (initialize).
 * nodes must have room for orc_pessimistic_n_nodes(n) entries; bb_indices for n.
 * Returns the number of nodes, or -1 on the unbucketable-centroid error, -2 when
 * all costs are NaN/inf so that no plane is chosen (reference: IndexError too). */
int64_t orc_initialize(const double *bb_coords, int64_t n, int n_buckets, int cells_per_leaf,
                       node_t *nodes, int64_t *bb_indices)
{
    for (int64_t i = 0; i < n; i++) bb_indices[i] = i;
    set_node(&nodes[0], 0, n, 0);
    int64_t node_index = 1;

    int64_t *temp = (int64_t *)malloc(sizeof(int64_t) * (n > 0 ? n : 1));
    bucket_t *buckets = (bucket_t *)malloc(sizeof(bucket_t) * n_buckets);
    /* double stack of (root_index, dim flag): utils.py:24-26, 52-68 */
    int64_t cap = INITIAL_STACK_LENGTH, size = 0;
    int64_t *stack = (int64_t *)malloc(sizeof(int64_t) * 2 * cap);
    stack[0] = 0; stack[1] = 0; size = 1;
    int64_t status = 0;

#define PUSH_BOTH(A, B) do { \
        if (size >= cap) { cap *= 2; stack = (int64_t *)realloc(stack, sizeof(int64_t) * 2 * cap); } \
        stack[2 * size] = (A); stack[2 * size + 1] = (B); size++; } while (0)

    while (size > 0) {
        size--;
        int64_t root_index = stack[2 * size];
        int64_t dim = stack[2 * size + 1];
        int64_t dim_flag = dim;
        if (dim < 0) dim += 2;

        node_t root = nodes[root_index];
        if (root.size <= cells_per_leaf) continue;

        double range_Rmin, range_Lmax;
        get_bounds(root.ptr, root.size, bb_coords, bb_indices, (int)dim, &range_Rmin, &range_Lmax);
        double bucket_length = (range_Lmax - range_Rmin) / (double)n_buckets;

        for (int i = 0; i < n_buckets; i++) {
            buckets[i].Max = (double)(i + 1) * bucket_length + range_Rmin;
            buckets[i].Min = (double)i * bucket_length + range_Rmin;
            buckets[i].Rmin = -1.0;
            buckets[i].Lmax = -1.0;
            buckets[i].index = -1;
            buckets[i].size = 0;
        }
        if (sort_bbox_indices(bb_indices, bb_coords, buckets, n_buckets, root.ptr, root.size, (int)dim, temp) != 0) {
            status = -1;
            break;
        }
        for (int i = 0; i < n_buckets; i++)
            get_bounds(buckets[i].index, buckets[i].size, bb_coords, bb_indices, (int)dim, &buckets[i].Rmin, &buckets[i].Lmax);

        if ((cells_per_leaf == 1) && (root.size == 2)) {                                  /* creation.py:312-320 */
            nodes[root_index].Lmax = range_Lmax;
            nodes[root_index].Rmin = range_Rmin;
            nodes[root_index].child = node_index;
            set_node(&nodes[node_index++], root.ptr, 1, !dim);
            set_node(&nodes[node_index++], root.ptr + 1, 1, !dim);
            continue;
        }

        /* drop / merge empty buckets: creation.py:322-341 (only Rmin/Lmax/index/size matter afterwards) */
        int nb = 0;
        for (int i = 0; i < n_buckets; i++)
            if (buckets[i].size != 0) buckets[nb++] = buckets[i];

        int needs_continue = 0;                                                           /* creation.py:345-358 */
        for (int i = 0; i < nb; i++) {
            if (buckets[i].size == root.size) {
                needs_continue = 1;
                if (dim_flag >= 0) {
                    dim_flag = (dim ? 0 : 1) - 2;
                    nodes[root_index].dim = (uint8_t)(root.dim ? 0 : 1);
                    PUSH_BOTH(root_index, dim_flag);
                } else {
                    nodes[root_index].Lmax = -1;
                    nodes[root_index].Rmin = -1;
                }
                break;
            }
        }
        if (needs_continue) continue;

        int plane;
        double Lmax, Rmin;
        split_plane(buckets, nb, root.size, range_Lmax, range_Rmin, bucket_length, &plane, &Lmax, &Rmin);
        if (plane >= nb) { status = -2; break; }
        int64_t right_index = buckets[plane].index;
        int64_t right_size = root.ptr + root.size - right_index;
        int64_t left_index = root.ptr;
        int64_t left_size = root.size - right_size;
        nodes[root_index].Lmax = Lmax;
        nodes[root_index].Rmin = Rmin;
        nodes[root_index].child = node_index;
        int64_t child_ind = node_index;
        set_node(&nodes[node_index++], left_index, left_size, !dim);
        set_node(&nodes[node_index++], right_index, right_size, !dim);
        PUSH_BOTH(child_ind + 1, dim ? 0 : 1);
        PUSH_BOTH(child_ind, dim ? 0 : 1);
    }
#undef PUSH_BOTH
    free(stack);
    free(buckets);
    free(temp);
    return status < 0 ? status : node_index;
}

/* ================================================================== */
/* Queries: query.py                                                   */
/* ================================================================== */
typedef struct { int64_t nodes_visited, cells_tested; } trav_stats;

static int64_t locate_point(const tree_t *t, pt point, double tolerance, trav_stats *st)  /* query.py:63-107 */
{
    istack stack;
    pt poly[MAX_N_VERTEX];
    stack_init(&stack);
    stack_push(&stack, 0);
    int64_t result = -1;
    const int M = (int)t->n_max_vert;
    while (stack.size > 0) {
        int64_t node_index = stack_pop(&stack);
        const node_t *node = &t->nodes[node_index];
        if (st) st->nodes_visited++;
        if (node->child == -1) {
            int found = 0;
            for (int64_t i = node->ptr; i < node->ptr + node->size; i++) {
                int64_t bbox_index = t->bb_indices[i];
                const int64_t *face = t->elements + bbox_index * M;
                int n = copy_vertices(t->vertices, face, M, poly);
                if (st) st->cells_tested++;
                if (point_in_polygon_or_on_edge(point, poly, n, tolerance)) { result = bbox_index; found = 1; break; }
            }
            if (found) break;
            continue;
        }
        int dim = node->dim ? 1 : 0;
        double pd = dim ? point.y : point.x;
        int left = pd <= node->Lmax;
        int right = pd >= node->Rmin;
        int64_t left_child = node->child;
        int64_t right_child = left_child + 1;
        if (left && right) {
            if ((node->Lmax - pd) < (pd - node->Rmin)) { stack_push(&stack, left_child); stack_push(&stack, right_child); }
            else { stack_push(&stack, right_child); stack_push(&stack, left_child); }
        } else if (left) stack_push(&stack, left_child);
        else if (right) stack_push(&stack, right_child);
    }
    stack_free(&stack);
    return result;
}

static void make_tree(tree_t *t, const int64_t *elements, int64_t n_elem, int64_t n_max_vert, const double *vertices,
                      const void *nodes, const int64_t *bb_indices, const double *bb_coords, const double *bbox)
{
    t->elements = elements; t->n_elem = n_elem; t->n_max_vert = n_max_vert; t->vertices = vertices;
    t->nodes = (const node_t *)nodes; t->bb_indices = bb_indices; t->bb_coords = bb_coords;
    memcpy(t->bbox, bbox, sizeof(double) * 4);
}

#define TREE_ARGS const int64_t *elements, int64_t n_elem, int64_t n_max_vert, const double *vertices, \
                  const void *nodes, const int64_t *bb_indices, const double *bb_coords, const double *bbox
#define TREE_PASS elements, n_elem, n_max_vert, vertices, nodes, bb_indices, bb_coords, bbox

/* query.py:110-117; stats (may be NULL) = {sum nodes visited, sum cells tested} */
void orc_locate_points(TREE_ARGS, const double *points, int64_t n_points, double tolerance, int64_t *result, int64_t *stats)
{
    tree_t t; make_tree(&t, TREE_PASS);
    int64_t nv = 0, ct = 0;
#pragma omp parallel for schedule(static) reduction(+ : nv, ct)
    for (int64_t i = 0; i < n_points; i++) {
        pt p = { points[2 * i], points[2 * i + 1] };
        trav_stats st = { 0, 0 };
        result[i] = locate_point(&t, p, tolerance, stats ? &st : NULL);
        nv += st.nodes_visited; ct += st.cells_tested;
    }
    if (stats) { stats[0] = nv; stats[1] = ct; }
}

static int64_t locate_point_on_edge(const tree_t *t, pt point, double tolerance)          /* query.py:121-165 */
{
    istack stack;
    stack_init(&stack);
    stack_push(&stack, 0);
    int64_t result = -1;
    while (stack.size > 0) {
        int64_t node_index = stack_pop(&stack);
        const node_t *node = &t->nodes[node_index];
        if (node->child == -1) {
            int found = 0;
            for (int64_t i = node->ptr; i < node->ptr + node->size; i++) {
                int64_t bbox_index = t->bb_indices[i];
                const int64_t *edge = t->elements + bbox_index * 2;
                pt v0 = { t->vertices[2 * edge[0]], t->vertices[2 * edge[0] + 1] };
                pt v1 = { t->vertices[2 * edge[1]], t->vertices[2 * edge[1] + 1] };
                if (point_on_edge(point, v0, v1, tolerance)) { result = bbox_index; found = 1; break; }
            }
            if (found) break;
            continue;
        }
        int dim = node->dim ? 1 : 0;
        double pd = dim ? point.y : point.x;
        int left = pd <= node->Lmax;
        int right = pd >= node->Rmin;
        int64_t left_child = node->child;
        int64_t right_child = left_child + 1;
        if (left && right) {
            if ((node->Lmax - pd) < (pd - node->Rmin)) { stack_push(&stack, left_child); stack_push(&stack, right_child); }
            else { stack_push(&stack, right_child); stack_push(&stack, left_child); }
        } else if (left) stack_push(&stack, left_child);
        else if (right) stack_push(&stack, right_child);
    }
    stack_free(&stack);
    return result;
}

void orc_locate_points_on_edge(TREE_ARGS, const double *points, int64_t n_points, double tolerance, int64_t *result) /* query.py:168-174 */
{
    tree_t t; make_tree(&t, TREE_PASS);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n_points; i++) {
        pt p = { points[2 * i], points[2 * i + 1] };
        result[i] = locate_point_on_edge(&t, p, tolerance);
    }
}

/* Growable per-chunk result buffers (query.py:247-275, 458-497): the reference
 * grows by doubling and retries the query; appending with realloc gives the
 * same final content. */
typedef struct { int64_t *ij; double *xy; int64_t size, cap; int with_xy; } pairbuf;

static void pairbuf_init(pairbuf *b, int64_t n, int with_xy)
{
    b->cap = n > 256 ? n : 256;
    b->size = 0;
    b->with_xy = with_xy;
    b->ij = (int64_t *)malloc(sizeof(int64_t) * 2 * b->cap);
    b->xy = with_xy ? (double *)malloc(sizeof(double) * 4 * b->cap) : NULL;
}
static inline void pairbuf_reserve(pairbuf *b)
{
    if (b->size >= b->cap) {
        b->cap *= 2;
        b->ij = (int64_t *)realloc(b->ij, sizeof(int64_t) * 2 * b->cap);
        if (b->with_xy) b->xy = (double *)realloc(b->xy, sizeof(double) * 4 * b->cap);
    }
}

static void locate_box(const tree_t *t, box_t box, pairbuf *out, int64_t index, trav_stats *st)  /* query.py:177-244 */
{
    if (!boxes_intersect(box, as_box(t->bbox))) return;
    istack stack;
    stack_init(&stack);
    stack_push(&stack, 0);
    while (stack.size > 0) {
        int64_t node_index = stack_pop(&stack);
        const node_t *node = &t->nodes[node_index];
        if (st) st->nodes_visited++;
        if (node->child == -1) {
            for (int64_t i = node->ptr; i < node->ptr + node->size; i++) {
                int64_t bbox_index = t->bb_indices[i];
                box_t leaf_box = as_box(t->bb_coords + 4 * bbox_index);
                if (st) st->cells_tested++;
                if (boxes_intersect(box, leaf_box)) {
                    pairbuf_reserve(out);
                    out->ij[2 * out->size] = index;
                    out->ij[2 * out->size + 1] = bbox_index;
                    out->size++;
                }
            }
        } else {
            int dim = node->dim ? 1 : 0;
            double bmin = dim ? box.ymin : box.xmin;
            double bmax = dim ? box.ymax : box.xmax;
            int left = bmin <= node->Lmax;
            int right = bmax >= node->Rmin;
            int64_t left_child = node->child;
            int64_t right_child = left_child + 1;
            if (left && right) { stack_push(&stack, left_child); stack_push(&stack, right_child); }
            else if (left) stack_push(&stack, left_child);
            else if (right) stack_push(&stack, right_child);
        }
    }
    stack_free(&stack);
}

static void chunk_bounds(int64_t n, int n_chunks, int c, int64_t *lo, int64_t *hi)       /* np.array_split, query.py:280-283 */
{
    int64_t base = n / n_chunks, rem = n % n_chunks;
    *lo = c * base + (c < rem ? c : rem);
    *hi = *lo + base + (c < rem ? 1 : 0);
}

static int default_chunks(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* Opaque result list handed back to Python: total count, then fetch. */
typedef struct { int n_chunks; pairbuf *chunks; int64_t total; } result_t;

int64_t orc_result_size(const result_t *r) { return r->total; }

void orc_result_fetch(result_t *r, int64_t *ii, int64_t *jj, double *xy)                 /* query.py:46-59, 524-533 */
{
    int64_t start = 0;
    for (int c = 0; c < r->n_chunks; c++) {
        pairbuf *b = &r->chunks[c];
        for (int64_t k = 0; k < b->size; k++) {
            ii[start + k] = b->ij[2 * k];
            jj[start + k] = b->ij[2 * k + 1];
        }
        if (xy && b->with_xy) memcpy(xy + 4 * start, b->xy, sizeof(double) * 4 * b->size);
        start += b->size;
    }
}

void orc_result_free(result_t *r)
{
    for (int c = 0; c < r->n_chunks; c++) { free(r->chunks[c].ij); free(r->chunks[c].xy); }
    free(r->chunks);
    free(r);
}

/* query.py:278-289 */
result_t *orc_locate_boxes(TREE_ARGS, const double *box_coords, int64_t n_box, int64_t *stats)
{
    tree_t t; make_tree(&t, TREE_PASS);
    int n_chunks = default_chunks();
    result_t *r = (result_t *)malloc(sizeof(result_t));
    r->n_chunks = n_chunks;
    r->chunks = (pairbuf *)malloc(sizeof(pairbuf) * n_chunks);
    int64_t nv = 0, ct = 0;
#pragma omp parallel for schedule(static, 1) reduction(+ : nv, ct)
    for (int c = 0; c < n_chunks; c++) {
        int64_t lo, hi;
        chunk_bounds(n_box, n_chunks, c, &lo, &hi);
        pairbuf_init(&r->chunks[c], hi - lo, 0);
        trav_stats st = { 0, 0 };
        for (int64_t i = lo; i < hi; i++) locate_box(&t, as_box(box_coords + 4 * i), &r->chunks[c], i, stats ? &st : NULL);
        nv += st.nodes_visited; ct += st.cells_tested;
    }
    r->total = 0;
    for (int c = 0; c < n_chunks; c++) r->total += r->chunks[c].size;
    if (stats) { stats[0] = nv; stats[1] = ct; }
    return r;
}

enum { INTERSECT_EDGE_EDGE = 0, INTERSECT_EDGE_FACE = 1 };                                /* query.py:331-335 */

static int compute_edge_edge_intersect(const tree_t *t, int64_t bbox_index, pt a, pt b, pt *c, pt *d) /* query.py:292-306 */
{
    const int64_t *tree_edge = t->elements + bbox_index * t->n_max_vert;
    pt p = { t->vertices[2 * tree_edge[0]], t->vertices[2 * tree_edge[0] + 1] };
    pt q = { t->vertices[2 * tree_edge[1]], t->vertices[2 * tree_edge[1] + 1] };
    double x, y;
    int intersects = lines_intersect(a, b, p, q, &x, &y);
    c->x = x; c->y = y;
    *d = *c;
    return intersects;
}

static int compute_edge_face_intersect(const tree_t *t, int64_t bbox_index, pt a, pt b, pt *c, pt *d) /* query.py:309-328 */
{
    box_t box = as_box(t->bb_coords + 4 * bbox_index);
    int intersects = cohen_sutherland_line_box_clip(a, b, box, c, d);
    if (intersects > 0) {
        pt polygon[MAX_N_VERTEX];
        int n = copy_vertices(t->vertices, t->elements + bbox_index * t->n_max_vert, (int)t->n_max_vert, polygon);
        double tolerance = nb_max(MIN_TOLERANCE, TOLERANCE_FACTOR * nb_max(box.xmax - box.xmin, box.ymax - box.ymin));
        intersects = cyrus_beck_line_polygon_clip(a, b, polygon, n, tolerance, c, d);
    }
    return intersects;
}

/* query.py:357-455; returns 0, or -1 on "Undefined clipping state" */
static int locate_edge(const tree_t *t, pt a, pt b, pairbuf *out, int64_t index, int intersect_type)
{
    pt c, d;
    int tree_intersects = cohen_sutherland_line_box_clip(a, b, as_box(t->bbox), &c, &d);
    if (tree_intersects < 0) return -1;
    if (!tree_intersects) return 0;
    pt V = to_vector(a, b);
    istack stack;
    stack_init(&stack);
    stack_push(&stack, 0);
    int status = 0;
    while (stack.size > 0) {
        int64_t node_index = stack_pop(&stack);
        const node_t *node = &t->nodes[node_index];
        if (node->child == -1) {
            for (int64_t i = node->ptr; i < node->ptr + node->size; i++) {
                int64_t bbox_index = t->bb_indices[i];
                int intersects = (intersect_type == INTERSECT_EDGE_EDGE)
                                     ? compute_edge_edge_intersect(t, bbox_index, a, b, &c, &d)
                                     : compute_edge_face_intersect(t, bbox_index, a, b, &c, &d);
                if (intersects < 0) { status = -1; intersects = 0; }
                if (intersects) {
                    pairbuf_reserve(out);
                    out->ij[2 * out->size] = index;
                    out->ij[2 * out->size + 1] = bbox_index;
                    out->xy[4 * out->size + 0] = c.x;
                    out->xy[4 * out->size + 1] = c.y;
                    out->xy[4 * out->size + 2] = d.x;
                    out->xy[4 * out->size + 3] = d.y;
                    out->size++;
                }
            }
            continue;
        }
        int node_dim = node->dim ? 1 : 0;
        double dx = node_dim ? V.y : V.x;
        double a_d = node_dim ? a.y : a.x;
        double b_d = node_dim ? b.y : b.x;
        double dx_left, dx_right;
        if (dx > 0.0) { dx_left = node->Lmax - a_d; dx_right = node->Rmin - b_d; }
        else { dx_left = node->Lmax - b_d; dx_right = node->Rmin - a_d; }
        int left = dx_left >= 0.0;
        int right = dx_right <= 0.0;
        if (dx > 0.0) {
            if (left) { double t_left = dx_left / dx; left = t_left >= 0.0; }
            if (right) { double t_right = dx_right / dx; right = t_right <= 1.0; }
        } else if (dx < 0.0) {
            if (left) { double t_left = 1.0 - (dx_left / dx); left = t_left >= 0.0; }
            if (right) { double t_right = 1.0 - (dx_right / dx); right = t_right <= 1.0; }
        }
        int64_t left_child = node->child;
        int64_t right_child = left_child + 1;
        if (left && right) { stack_push(&stack, left_child); stack_push(&stack, right_child); }
        else if (left) stack_push(&stack, left_child);
        else if (right) stack_push(&stack, right_child);
    }
    stack_free(&stack);
    return status;
}

/* query.py:500-544; *status = -1 on "Undefined clipping state" */
result_t *orc_locate_edges(TREE_ARGS, const double *edge_coords, int64_t n_edge, int intersect_type, int *status)
{
    tree_t t; make_tree(&t, TREE_PASS);
    int n_chunks = default_chunks();
    result_t *r = (result_t *)malloc(sizeof(result_t));
    r->n_chunks = n_chunks;
    r->chunks = (pairbuf *)malloc(sizeof(pairbuf) * n_chunks);
    int st = 0;
#pragma omp parallel for schedule(static, 1) reduction(min : st)
    for (int c = 0; c < n_chunks; c++) {
        int64_t lo, hi;
        chunk_bounds(n_edge, n_chunks, c, &lo, &hi);
        pairbuf_init(&r->chunks[c], hi - lo, 1);
        for (int64_t i = lo; i < hi; i++) {
            pt a = { edge_coords[4 * i], edge_coords[4 * i + 1] };
            pt b = { edge_coords[4 * i + 2], edge_coords[4 * i + 3] };
            int s = locate_edge(&t, a, b, &r->chunks[c], i, intersect_type);
            if (s < st) st = s;
        }
    }
    r->total = 0;
    for (int c = 0; c < n_chunks; c++) r->total += r->chunks[c].size;
    if (status) *status = st;
    return r;
}

/* ================================================================== */
/* Pair kernels: algorithms/                                           */
/* ================================================================== */
/* sutherland_hodgman.py:151-168 */
void orc_area_of_intersection(const double *vertices_a, const double *vertices_b, const int64_t *faces_a, int ma,
                              const int64_t *faces_b, int mb, const int64_t *indices_a, const int64_t *indices_b,
                              int64_t n, double *area)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) {
        pt a[MAX_N_VERTEX], b[MAX_N_VERTEX];
        int la = copy_vertices(vertices_a, faces_a + indices_a[i] * ma, ma, a);
        int lb = copy_vertices(vertices_b, faces_b + indices_b[i] * mb, mb, b);
        area[i] = polygon_polygon_clip_area(a, la, b, lb);
    }
}

/* sutherland_hodgman.py:171-187 */
void orc_box_area_of_intersection(const double *bbox_coords, const double *vertices, const int64_t *faces, int m,
                                  const int64_t *indices_bbox, const int64_t *indices_face, int64_t n, double *area)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) {
        pt a[4], b[MAX_N_VERTEX];
        copy_box_vertices(as_box(bbox_coords + 4 * indices_bbox[i]), a);
        int lb = copy_vertices(vertices, faces + indices_face[i] * m, m, b);
        area[i] = polygon_polygon_clip_area(a, 4, b, lb);
    }
}

/* separating_axis.py:58-75 */
void orc_polygons_intersect(const double *vertices_a, const double *vertices_b, const int64_t *faces_a, int ma,
                            const int64_t *faces_b, int mb, const int64_t *indices_a, const int64_t *indices_b,
                            int64_t n, uint8_t *intersects)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) {
        pt a[MAX_N_VERTEX], b[MAX_N_VERTEX];
        int la = copy_vertices(vertices_a, faces_a + indices_a[i] * ma, ma, a);
        int lb = copy_vertices(vertices_b, faces_b + indices_b[i] * mb, mb, b);
        intersects[i] = (uint8_t)(separating_axes(a, la, b, lb) && separating_axes(b, lb, a, la));
    }
}

/* barycentric_triangle.py:46-64; weights (n,3) pre-zeroed by the caller */
void orc_barycentric_triangle_weights(const double *points, const int64_t *face_indices, const int64_t *faces, int m,
                                      const double *vertices, int64_t n, double *weights)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) {
        int64_t face_index = face_indices[i];
        if (face_index == -1) continue;
        const int64_t *face = faces + face_index * m;
        pt a = { vertices[2 * face[0]], vertices[2 * face[0] + 1] };
        pt b = { vertices[2 * face[1]], vertices[2 * face[1] + 1] };
        pt c = { vertices[2 * face[2]], vertices[2 * face[2] + 1] };
        pt p = { points[2 * i], points[2 * i + 1] };
        tri_compute_weights(a, b, c, p, weights + 3 * i);
    }
}

/* barycentric_wachspress.py:88-107; weights (n,m) pre-zeroed by the caller */
void orc_barycentric_wachspress_weights(const double *points, const int64_t *face_indices, const int64_t *faces, int m,
                                        const double *vertices, double tolerance, int64_t n, double *weights)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) {
        int64_t face_index = face_indices[i];
        if (face_index == -1) continue;
        pt polygon[MAX_N_VERTEX];
        int len = copy_vertices(vertices, faces + face_index * m, m, polygon);
        pt p = { points[2 * i], points[2 * i + 1] };
        wachspress_compute_weights(polygon, len, p, weights + (int64_t)m * i, m, tolerance);
    }
}

int orc_num_threads(void) { return default_chunks(); }
void orc_set_num_threads(int n)
{
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}
