// algorithms.cu -- batched C-ABI entries for the geometry helpers the reference exports beside its query API
// (numba_celltree.algorithms.__all__ and geometry_utils).  The reference's versions are scalar
// @njit functions on Point / Box tuples; here one call handles n inputs, one thread each, with the same fp64 device
// functions the query kernels use (geometry.cuh), so the bits are the reference's.
#include "geometry.cuh"

namespace ct {

CT_DEV bool point_inside_box(P2 a, const Box4 &box) {  // geometry_utils.py:527-529
    return box.xmin < a.x && a.x < box.xmax && box.ymin < a.y && a.y < box.ymax;
}

// algorithms/liang_barsky.py:10-64.  The four box sides are taken as (P, Q) pairs in the order left, right, lower,
// upper; a side parallel to the segment (P == 0) rejects when the segment lies outside of it (Q < 0), any other side
// moves t0 up (entering, P < 0) or t1 down (leaving, P > 0).  Zero-length segments and t0 == t1 (touching) miss.
CT_DEV bool liang_barsky_line_box_clip(P2 a, P2 b, const Box4 &box, P2 &c, P2 &d) {
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    c = P2{nan, nan};
    d = P2{nan, nan};
    const double dx = b.x - a.x, dy = b.y - a.y;
    if (dx == 0.0 && dy == 0.0) return false;
    if (point_inside_box(a, box) && point_inside_box(b, box)) {
        c = a;
        d = b;
        return true;
    }
    double t0 = 0.0, t1 = 1.0;
#pragma unroll
    for (int side = 0; side < 4; side++) {
        const double p = side == 0 ? -dx : (side == 1 ? dx : (side == 2 ? -dy : dy));
        const double q = side == 0 ? a.x - box.xmin : (side == 1 ? box.xmax - a.x : (side == 2 ? a.y - box.ymin : box.ymax - a.y));
        if (p == 0.0) {
            if (q < 0.0) return false;
        } else {
            const double t = q / p;
            if (p < 0.0) {
                if (t > t1) return false;
                else if (t > t0) t0 = t;
            } else if (p > 0.0) {
                if (t < t0) return false;
                else if (t < t1) t1 = t;
            }
        }
    }
    if (t0 == t1) return false;
    c = P2{a.x + t0 * dx, a.y + t0 * dy};
    d = P2{a.x + t1 * dx, a.y + t1 * dy};
    return true;
}

// geometry_utils.py:98-147: the plain crossing-number test; no tolerance, no on-edge acceptance, every edge counts
// (also zero-length ones: their y-test is false)
CT_DEV bool point_in_polygon(P2 p, const double2 *__restrict__ poly, int length) {
    double2 v0 = poly[length - 1];
    bool inside = false;
    for (int i = 0; i < length; i++) {
        const double2 v1 = poly[i];
        if (((v0.y > p.y) != (v1.y > p.y)) && (p.x < ((v1.x - v0.x) * (p.y - v0.y) / (v1.y - v0.y) + v0.x))) inside = !inside;
        v0 = v1;
    }
    return inside;
}

// geometry_utils.py:241-270
CT_DEV bool point_in_triangle(P2 p, P2 ta, P2 tb, P2 tc, double tolerance) {
    const P2 ap = to_vector(ta, p), bp = to_vector(tb, p), cp = to_vector(tc, p);
    const P2 ab = to_vector(ta, tb), bc = to_vector(tb, tc), ca = to_vector(tc, ta);
    const double A = cross_product(ab, ap), B = cross_product(bc, bp), C = cross_product(ca, cp);
    const bool sA = A > 0, sB = B > 0, sC = C > 0;
    if (sA == sB && sB == sC) return true;
    return (within_perpendicular_distance(A, ab, tolerance) && in_bounds(p, ta, tb)) ||
           (within_perpendicular_distance(B, bc, tolerance) && in_bounds(p, tb, tc)) ||
           (within_perpendicular_distance(C, ca, tolerance) && in_bounds(p, tc, ta));
}

enum { CLIP_COHEN_SUTHERLAND = 0, CLIP_LIANG_BARSKY = 1 };

template <int WHICH>
__global__ void __launch_bounds__(256) k_line_box_clip(const double2 *__restrict__ a, const double2 *__restrict__ b,
                                                        const double *__restrict__ boxes, int64_t box_stride, int64_t n,
                                                        uint8_t *__restrict__ hit, double2 *__restrict__ c, double2 *__restrict__ d) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const double2 pa = a[i], pb = b[i];
    const Box4 box = load_box(boxes, i * box_stride);
    P2 pc, pd;
    bool ok;
    if (WHICH == CLIP_LIANG_BARSKY) ok = liang_barsky_line_box_clip(P2{pa.x, pa.y}, P2{pb.x, pb.y}, box, pc, pd);
    else ok = cohen_sutherland_line_box_clip(P2{pa.x, pa.y}, P2{pb.x, pb.y}, box, pc, pd) != 0;
    hit[i] = ok ? 1 : 0;
    c[i] = make_double2(pc.x, pc.y);
    d[i] = make_double2(pd.x, pd.y);
}

__global__ void __launch_bounds__(128) k_line_polygon_clip(const double2 *__restrict__ a, const double2 *__restrict__ b, int64_t n,
                                                            const double2 *__restrict__ poly, int length, double tolerance,
                                                            uint8_t *__restrict__ hit, double2 *__restrict__ c, double2 *__restrict__ d) {
    const int64_t i = (int64_t)blockIdx.x * 128 + threadIdx.x;
    if (i >= n) return;
    Poly<MAX_N_VERTEX> polygon;
    polygon.n = length;
    for (int k = 0; k < length; k++) {
        polygon.x[k] = poly[k].x;
        polygon.y[k] = poly[k].y;
    }
    const double2 pa = a[i], pb = b[i];
    P2 pc, pd;
    const bool ok = cyrus_beck_line_polygon_clip<MAX_N_VERTEX>(P2{pa.x, pa.y}, P2{pb.x, pb.y}, polygon, tolerance, pc, pd);
    hit[i] = ok ? 1 : 0;
    c[i] = make_double2(pc.x, pc.y);
    d[i] = make_double2(pd.x, pd.y);
}

__global__ void __launch_bounds__(256) k_points_in_polygon(const double2 *__restrict__ points, int64_t n, const double2 *__restrict__ poly,
                                                            int length, uint8_t *__restrict__ inside) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const double2 p = points[i];
    inside[i] = point_in_polygon(P2{p.x, p.y}, poly, length) ? 1 : 0;
}

__global__ void __launch_bounds__(256) k_points_in_triangles(const double2 *__restrict__ points, const int64_t *__restrict__ face_indices,
                                                              int64_t n, const int64_t *__restrict__ faces, int64_t n_face, int n_max_vert,
                                                              const double2 *__restrict__ vertices, int64_t n_vertex, double tolerance,
                                                              uint8_t *__restrict__ inside, int *__restrict__ bad) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const int64_t f = face_indices[i];
    if (f < 0 || f >= n_face) {  // the reference would index out of bounds here
        *bad = 1;
        inside[i] = 0;
        return;
    }
    const int64_t *face = faces + f * n_max_vert;
    const int64_t i0 = face[0], i1 = face[1], i2 = face[2];
    if (i0 < 0 || i1 < 0 || i2 < 0 || i0 >= n_vertex || i1 >= n_vertex || i2 >= n_vertex) {
        *bad = 1;
        inside[i] = 0;
        return;
    }
    const double2 p = points[i], ta = vertices[i0], tb = vertices[i1], tc = vertices[i2];
    inside[i] = point_in_triangle(P2{p.x, p.y}, P2{ta.x, ta.y}, P2{tb.x, tb.y}, P2{tc.x, tc.y}, tolerance) ? 1 : 0;
}

static int line_box_clip(int which, const double *a, const double *b, const double *boxes, int64_t n_boxes, int64_t n, uint8_t *hit,
                         double *c, double *d, int32_t mem) {
    if (n < 0 || (n > 0 && (!a || !b || !boxes || !hit || !c || !d)) || (n_boxes != 1 && n_boxes != n)) {
        set_error("line / box clip: null argument, or n_boxes is neither 1 nor n");
        return CT_ERR_VALUE;
    }
    if (n == 0) return CT_OK;
    cudaStream_t s = current_stream();
    DevIn<double> da, db, dbox;
    DevOut<uint8_t> dhit;
    DevOut<double> dc, dd;
    CT_CHECK(da.init(a, 2 * n, mem, s));
    CT_CHECK(db.init(b, 2 * n, mem, s));
    CT_CHECK(dbox.init(boxes, 4 * n_boxes, mem, s));
    CT_CHECK(dhit.init(hit, n, mem, s));
    CT_CHECK(dc.init(c, 2 * n, mem, s));
    CT_CHECK(dd.init(d, 2 * n, mem, s));
    const int64_t stride = n_boxes == 1 ? 0 : 1;
    auto pa = reinterpret_cast<const double2 *>(da.p), pb = reinterpret_cast<const double2 *>(db.p);
    auto pc = reinterpret_cast<double2 *>(dc.p), pd = reinterpret_cast<double2 *>(dd.p);
    if (which == CLIP_LIANG_BARSKY)
        k_line_box_clip<CLIP_LIANG_BARSKY><<<grid_for(n, 256), 256, 0, s>>>(pa, pb, dbox.p, stride, n, dhit.p, pc, pd);
    else
        k_line_box_clip<CLIP_COHEN_SUTHERLAND><<<grid_for(n, 256), 256, 0, s>>>(pa, pb, dbox.p, stride, n, dhit.p, pc, pd);
    CT_LAUNCH_CHECK();
    CT_CHECK(dhit.finish(s));
    CT_CHECK(dc.finish(s));
    CT_CHECK(dd.finish(s));
    if (mem != CT_MEM_DEVICE) CT_CUDA(cudaStreamSynchronize(s));
    return CT_OK;
}

}  // namespace ct

using namespace ct;

extern "C" int ct_liang_barsky_line_box_clip(const double *a, const double *b, const double *boxes, int64_t n_boxes, int64_t n,
                                             uint8_t *intersects, double *c, double *d, int32_t mem) {
    return line_box_clip(CLIP_LIANG_BARSKY, a, b, boxes, n_boxes, n, intersects, c, d, mem);
}

extern "C" int ct_cohen_sutherland_line_box_clip(const double *a, const double *b, const double *boxes, int64_t n_boxes, int64_t n,
                                                 uint8_t *intersects, double *c, double *d, int32_t mem) {
    return line_box_clip(CLIP_COHEN_SUTHERLAND, a, b, boxes, n_boxes, n, intersects, c, d, mem);
}

extern "C" int ct_cyrus_beck_line_polygon_clip(const double *a, const double *b, int64_t n, const double *polygon, int32_t n_polygon,
                                               double tolerance, uint8_t *intersects, double *c, double *d, int32_t mem) {
    if (n < 0 || n_polygon < 3 || n_polygon > MAX_N_VERTEX || !polygon || (n > 0 && (!a || !b || !intersects || !c || !d))) {
        set_error("ct_cyrus_beck_line_polygon_clip: null argument, or a polygon of fewer than 3 / more than 32 vertices");
        return CT_ERR_VALUE;
    }
    if (n == 0) return CT_OK;
    cudaStream_t s = current_stream();
    DevIn<double> da, db, dpoly;
    DevOut<uint8_t> dhit;
    DevOut<double> dc, dd;
    CT_CHECK(da.init(a, 2 * n, mem, s));
    CT_CHECK(db.init(b, 2 * n, mem, s));
    CT_CHECK(dpoly.init(polygon, 2 * (size_t)n_polygon, mem, s));
    CT_CHECK(dhit.init(intersects, n, mem, s));
    CT_CHECK(dc.init(c, 2 * n, mem, s));
    CT_CHECK(dd.init(d, 2 * n, mem, s));
    k_line_polygon_clip<<<grid_for(n, 128), 128, 0, s>>>(reinterpret_cast<const double2 *>(da.p), reinterpret_cast<const double2 *>(db.p), n,
                                                        reinterpret_cast<const double2 *>(dpoly.p), n_polygon, tolerance, dhit.p,
                                                        reinterpret_cast<double2 *>(dc.p), reinterpret_cast<double2 *>(dd.p));
    CT_LAUNCH_CHECK();
    CT_CHECK(dhit.finish(s));
    CT_CHECK(dc.finish(s));
    CT_CHECK(dd.finish(s));
    if (mem != CT_MEM_DEVICE) CT_CUDA(cudaStreamSynchronize(s));
    return CT_OK;
}

extern "C" int ct_points_in_polygon(const double *points, int64_t n, const double *polygon, int32_t n_polygon, uint8_t *inside, int32_t mem) {
    if (n < 0 || n_polygon < 1 || !polygon || (n > 0 && (!points || !inside))) {
        set_error("ct_points_in_polygon: null argument or empty polygon");
        return CT_ERR_VALUE;
    }
    if (n == 0) return CT_OK;
    cudaStream_t s = current_stream();
    DevIn<double> dp, dpoly;
    DevOut<uint8_t> dout;
    CT_CHECK(dp.init(points, 2 * n, mem, s));
    CT_CHECK(dpoly.init(polygon, 2 * (size_t)n_polygon, mem, s));
    CT_CHECK(dout.init(inside, n, mem, s));
    k_points_in_polygon<<<grid_for(n, 256), 256, 0, s>>>(reinterpret_cast<const double2 *>(dp.p), n, reinterpret_cast<const double2 *>(dpoly.p),
                                                        n_polygon, dout.p);
    CT_LAUNCH_CHECK();
    CT_CHECK(dout.finish(s));
    if (mem != CT_MEM_DEVICE) CT_CUDA(cudaStreamSynchronize(s));
    return CT_OK;
}

extern "C" int ct_points_in_triangles(const double *points, const int64_t *face_indices, int64_t n, const int64_t *faces, int64_t n_face,
                                      int32_t n_max_vert, const double *vertices, int64_t n_vertex, double tolerance, uint8_t *inside,
                                      int32_t mem) {
    if (n < 0 || n_max_vert < 3 || (n > 0 && (!points || !face_indices || !faces || !vertices || !inside))) {
        set_error("ct_points_in_triangles: null argument or faces with fewer than 3 columns");
        return CT_ERR_VALUE;
    }
    if (n == 0) return CT_OK;
    cudaStream_t s = current_stream();
    DevIn<double> dp, dv;
    DevIn<int64_t> dfi, df;
    DevOut<uint8_t> dout;
    Scratch<int> bad;
    CT_CHECK(dp.init(points, 2 * n, mem, s));
    CT_CHECK(dfi.init(face_indices, n, mem, s));
    CT_CHECK(df.init(faces, (size_t)n_face * n_max_vert, mem, s));
    CT_CHECK(dv.init(vertices, 2 * (size_t)n_vertex, mem, s));
    CT_CHECK(dout.init(inside, n, mem, s));
    CT_CHECK(bad.alloc(1, s));
    CT_CUDA(cudaMemsetAsync(bad.p, 0, sizeof(int), s));
    k_points_in_triangles<<<grid_for(n, 256), 256, 0, s>>>(reinterpret_cast<const double2 *>(dp.p), dfi.p, n, df.p, n_face, n_max_vert,
                                                          reinterpret_cast<const double2 *>(dv.p), n_vertex, tolerance, dout.p, bad.p);
    CT_LAUNCH_CHECK();
    CT_CHECK(dout.finish(s));
    int host_bad = 0;
    CT_CUDA(cudaMemcpyAsync(&host_bad, bad.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    CT_CUDA(cudaStreamSynchronize(s));
    if (host_bad) {
        set_error("ct_points_in_triangles: a face index or vertex index is out of range");
        return CT_ERR_VALUE;
    }
    return CT_OK;
}
