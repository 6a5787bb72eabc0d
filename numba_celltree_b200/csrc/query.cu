// query.cu -- query kernels and their C-ABI entry points (include/celltree_b200.h).
//
//   ct_locate_points     one thread per point; fused barycentric weights on the hit polygon
//   ct_locate_boxes      count kernel -> exclusive scan (CUB) -> fill kernel [-> clip area -> compact]
//   ct_locate_faces      ccw + bbox of the query faces -> box count/scan/fill -> SAT filter -> compact
//                        [-> Sutherland-Hodgman area -> compact]
//   ct_intersect_edges   count -> scan -> fill -> per-edge stable sort by t
//
// Variable-length results keep the reference's order: query index ascending (the scan runs over the
// counts in query order) and, within a query, DFS emission order (one thread walks one query in both
// passes, so count and fill take identical decisions).
#include <cub/device/device_scan.cuh>
#include <cub/iterator/counting_input_iterator.cuh>
#include <cub/iterator/transform_input_iterator.cuh>

#include "traverse.cuh"

namespace ct {

// ---- library state -----------------------------------------------------------------------------------
static thread_local std::string g_error;
static thread_local cudaStream_t g_stream = 0;
static int64_t g_launches = 0;

void set_error(const std::string &msg) { g_error = msg; }
cudaStream_t current_stream() { return g_stream; }
void count_launch(int n) { __atomic_fetch_add(&g_launches, (int64_t)n, __ATOMIC_RELAXED); }

constexpr int BLOCK = 128;

// Device view of a caller array: for CT_MEM_HOST a stream-ordered scratch copy, else the pointer itself.
template <typename T>
struct DevIn {
    Scratch<T> owned;
    const T *p = nullptr;
    int init(const T *src, size_t count, int mem, cudaStream_t s) {
        if (mem == CT_MEM_DEVICE) {
            p = src;
            return CT_OK;
        }
        CT_CHECK(owned.alloc(count, s));
        if (count) CT_CUDA(cudaMemcpyAsync(owned.p, src, count * sizeof(T), cudaMemcpyHostToDevice, s));
        p = owned.p;
        return CT_OK;
    }
};

template <typename T>
struct DevOut {
    Scratch<T> owned;
    T *p = nullptr;
    T *host = nullptr;
    size_t count = 0;
    int init(T *dst, size_t n, int mem, cudaStream_t s) {
        count = n;
        if (dst == nullptr) return CT_OK;
        if (mem == CT_MEM_DEVICE) {
            p = dst;
            return CT_OK;
        }
        host = dst;
        CT_CHECK(owned.alloc(n, s));
        p = owned.p;
        return CT_OK;
    }
    int finish(cudaStream_t s) {
        if (host && count) CT_CUDA(cudaMemcpyAsync(host, p, count * sizeof(T), cudaMemcpyDeviceToHost, s));
        return CT_OK;
    }
};

// ---- point kernels -------------------------------------------------------------------------------------
template <int MAXV, bool WEIGHTS>
__global__ void __launch_bounds__(BLOCK) k_locate_points(TreeView t, const double2 *__restrict__ points, int64_t n,
                                                         double tolerance, int64_t *__restrict__ out,
                                                         double *__restrict__ weights) {
    int64_t i = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
    if (i >= n) return;
    double2 pt = __ldg(points + i);
    P2 p{pt.x, pt.y};
    Poly<MAXV> poly;
    int found = locate_point<MAXV>(t, p, tolerance, poly);
    out[i] = found;
    if constexpr (WEIGHTS) {
        const int M = t.M;
        double *w_out = weights + i * (int64_t)M;
        if constexpr (MAXV == 3) {
            // barycentric_triangle_weights, algorithms/barycentric_triangle.py:46-64
            double u = 0.0, v = 0.0, w = 0.0;
            if (found != -1)
                triangle_weights(P2{poly.x[0], poly.y[0]}, P2{poly.x[1], poly.y[1]}, P2{poly.x[2], poly.y[2]}, p, u, v, w);
            w_out[0] = u;
            w_out[1] = v;
            w_out[2] = w;
        } else {
            // barycentric_wachspress_weights, algorithms/barycentric_wachspress.py:88-107
            double w[MAXV];
#pragma unroll
            for (int k = 0; k < MAXV; k++) w[k] = 0.0;
            if (found != -1) wachspress_weights<MAXV>(poly, p, tolerance, w);
            if constexpr (MAXV == 4) {
                if (M == 4) {
                    double2 *o = reinterpret_cast<double2 *>(w_out);
                    o[0] = make_double2(w[0], w[1]);
                    o[1] = make_double2(w[2], w[3]);
                    return;
                }
            }
#pragma unroll
            for (int k = 0; k < MAXV; k++)
                if (k < M) w_out[k] = w[k];
        }
    }
}

__global__ void __launch_bounds__(BLOCK) k_locate_points_on_edge(TreeView t, const double2 *__restrict__ points, int64_t n,
                                                                 double tolerance, int64_t *__restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
    if (i >= n) return;
    double2 pt = __ldg(points + i);
    out[i] = locate_point_on_edge(t, P2{pt.x, pt.y}, tolerance);
}

template <int MAXV>
static int launch_locate_points(const TreeView &v, const double2 *pts, int64_t n, double tol, int64_t *out, double *weights,
                                cudaStream_t s) {
    int grid = grid_for(n, BLOCK);
    if (weights)
        k_locate_points<MAXV, true><<<grid, BLOCK, 0, s>>>(v, pts, n, tol, out, weights);
    else
        k_locate_points<MAXV, false><<<grid, BLOCK, 0, s>>>(v, pts, n, tol, out, weights);
    CT_LAUNCH_CHECK();
    return CT_OK;
}

static int check_depth(const ct_tree *tree) {
    if (tree->depth > STACK_CAP) {
        char buf[160];
        snprintf(buf, sizeof(buf), "tree has %d levels; the traversal stack holds %d", tree->depth, STACK_CAP);
        set_error(buf);
        return CT_ERR_DEPTH;
    }
    return CT_OK;
}

// Host pointers: pipeline chunks over two streams so that H2D, traversal and D2H overlap.
static int locate_points_device(const ct_tree *tree, const double2 *pts, int64_t n, double tol, int64_t *out, double *weights,
                                cudaStream_t s) {
    if (n == 0) return CT_OK;
    TreeView v = tree->view();
    if (tree->kind == CT_KIND_EDGES) {
        k_locate_points_on_edge<<<grid_for(n, BLOCK), BLOCK, 0, s>>>(v, pts, n, tol, out);
        CT_LAUNCH_CHECK();
        return CT_OK;
    }
    if (tree->M == 3) return launch_locate_points<3>(v, pts, n, tol, out, weights, s);
    if (tree->M == 4) return launch_locate_points<4>(v, pts, n, tol, out, weights, s);
    if (tree->M <= 8) return launch_locate_points<8>(v, pts, n, tol, out, weights, s);
    return launch_locate_points<32>(v, pts, n, tol, out, weights, s);
}

}  // namespace ct

using namespace ct;

extern "C" int ct_locate_points(const ct_tree *tree, const double *points, int64_t n, double tolerance, int64_t *out_index,
                                double *weights, int32_t mem) {
    if (!tree || n < 0 || (n > 0 && (!points || !out_index))) {
        set_error("ct_locate_points: null argument");
        return CT_ERR_VALUE;
    }
    if (weights && tree->kind != CT_KIND_FACES) {
        set_error("ct_locate_points: barycentric weights need a face tree");
        return CT_ERR_VALUE;
    }
    CT_CHECK(check_depth(tree));
    CT_CUDA(cudaSetDevice(tree->device));
    cudaStream_t s = current_stream();
    if (mem == CT_MEM_DEVICE)
        return locate_points_device(tree, reinterpret_cast<const double2 *>(points), n, tolerance, out_index, weights, s);

    // host buffers: chunked pipeline on two private streams (copy-in / kernel / copy-out overlap)
    const int64_t CHUNK = 1 << 22;
    const int NS = 2;
    const int M = tree->M;
    cudaStream_t st[NS];
    double2 *d_pts[NS] = {nullptr, nullptr};
    int64_t *d_out[NS] = {nullptr, nullptr};
    double *d_w[NS] = {nullptr, nullptr};
    const int64_t chunk = n < CHUNK ? (n > 0 ? n : 1) : CHUNK;
    int status = CT_OK;
    CT_CUDA(cudaStreamSynchronize(s));
    for (int k = 0; k < NS; k++) CT_CUDA(cudaStreamCreateWithFlags(&st[k], cudaStreamNonBlocking));
    auto body = [&]() -> int {
        for (int k = 0; k < NS; k++) {
            CT_CHECK(dalloc(&d_pts[k], chunk, st[k]));
            CT_CHECK(dalloc(&d_out[k], chunk, st[k]));
            if (weights) CT_CHECK(dalloc(&d_w[k], chunk * M, st[k]));
        }
        int k = 0;
        for (int64_t lo = 0; lo < n; lo += chunk, k = (k + 1) % NS) {
            int64_t m = (n - lo) < chunk ? (n - lo) : chunk;
            CT_CUDA(cudaMemcpyAsync(d_pts[k], points + 2 * lo, m * sizeof(double2), cudaMemcpyHostToDevice, st[k]));
            CT_CHECK(locate_points_device(tree, d_pts[k], m, tolerance, d_out[k], weights ? d_w[k] : nullptr, st[k]));
            CT_CUDA(cudaMemcpyAsync(out_index + lo, d_out[k], m * sizeof(int64_t), cudaMemcpyDeviceToHost, st[k]));
            if (weights)
                CT_CUDA(cudaMemcpyAsync(weights + lo * M, d_w[k], m * M * sizeof(double), cudaMemcpyDeviceToHost, st[k]));
        }
        for (int q = 0; q < NS; q++) CT_CUDA(cudaStreamSynchronize(st[q]));
        return CT_OK;
    };
    status = body();
    for (int k = 0; k < NS; k++) {
        dfree(d_pts[k], st[k]);
        dfree(d_out[k], st[k]);
        dfree(d_w[k], st[k]);
        cudaStreamSynchronize(st[k]);
        cudaStreamDestroy(st[k]);
    }
    return status;
}

namespace ct {

// ---- box kernels ---------------------------------------------------------------------------------------------
template <bool FILL>
__global__ void __launch_bounds__(BLOCK) k_locate_boxes(TreeView t, const double *__restrict__ boxes, int64_t n,
                                                        int32_t *__restrict__ counts, const int64_t *__restrict__ offsets,
                                                        int32_t *__restrict__ out_i, int32_t *__restrict__ out_j) {
    int64_t q = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
    if (q >= n) return;
    Box4 box = load_box(boxes, q);
    if constexpr (FILL) {
        int64_t base = offsets[q];
        locate_box(t, box, [&](int k, int bbox_index) {
            out_i[base + k] = (int32_t)q;
            out_j[base + k] = bbox_index;
        });
    } else {
        counts[q] = locate_box(t, box, [](int, int) {});
    }
}

// ---- edge kernels ----------------------------------------------------------------------------------------------
template <int MAXV, bool FILL>
__global__ void __launch_bounds__(BLOCK) k_locate_edges(TreeView t, const double *__restrict__ edges, int64_t n,
                                                        int32_t *__restrict__ counts, const int64_t *__restrict__ offsets,
                                                        int32_t *__restrict__ out_i, int32_t *__restrict__ out_j,
                                                        double *__restrict__ out_xy) {
    int64_t q = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
    if (q >= n) return;
    const double2 *e = reinterpret_cast<const double2 *>(edges) + 2 * q;
    double2 a2 = __ldg(e), b2 = __ldg(e + 1);
    P2 a{a2.x, a2.y}, b{b2.x, b2.y};
    if constexpr (FILL) {
        int64_t base = offsets[q];
        locate_edge<MAXV>(t, a, b, [&](int k, int bbox_index, P2 c, P2 d) {
            out_i[base + k] = (int32_t)q;
            out_j[base + k] = bbox_index;
            double2 *o = reinterpret_cast<double2 *>(out_xy + 4 * (base + k));
            o[0] = make_double2(c.x, c.y);
            o[1] = make_double2(d.x, d.y);
        });
    } else {
        counts[q] = locate_edge<MAXV>(t, a, b, [](int, int, P2, P2) {});
    }
}

// sort_intersections_by_edge, geometry_utils.py:564-574: within each query edge's (already contiguous)
// range, stable sort by t = (c - a) . (b - a); np.lexsort puts NaN last and keeps ties in input order.
__global__ void __launch_bounds__(BLOCK) k_sort_edge_ranges(const double *__restrict__ edges, int64_t n,
                                                            const int64_t *__restrict__ offsets, int32_t *__restrict__ out_j,
                                                            double *__restrict__ out_xy) {
    int64_t q = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
    if (q >= n) return;
    int64_t lo = offsets[q], hi = offsets[q + 1];
    if (hi - lo < 2) return;
    const double2 *e = reinterpret_cast<const double2 *>(edges) + 2 * q;
    double2 a = __ldg(e), b = __ldg(e + 1);
    double abx = b.x - a.x, aby = b.y - a.y;
    double2 *xy = reinterpret_cast<double2 *>(out_xy);
    auto t_of = [&](double2 c) { return (c.x - a.x) * abx + (c.y - a.y) * aby; };
    auto lt = [](double x, double y) { return x < y || (y != y && x == x); };
    for (int64_t k = lo + 1; k < hi; k++) {
        double2 c = xy[2 * k], d = xy[2 * k + 1];
        int32_t j = out_j[k];
        double tk = t_of(c);
        int64_t m = k;
        while (m > lo) {
            double2 cp = xy[2 * (m - 1)];
            if (!lt(tk, t_of(cp))) break;
            xy[2 * m] = cp;
            xy[2 * m + 1] = xy[2 * (m - 1) + 1];
            out_j[m] = out_j[m - 1];
            m--;
        }
        if (m != k) {
            xy[2 * m] = c;
            xy[2 * m + 1] = d;
            out_j[m] = j;
        }
    }
}

// ---- pair kernels ------------------------------------------------------------------------------------------------
CT_DEV void box_polygon(const Box4 &box, Poly<4> &a) {  // copy_box_vertices, geometry_utils.py:513-524
    a.n = 4;
    a.x[0] = box.xmin; a.y[0] = box.ymin;
    a.x[1] = box.xmax; a.y[1] = box.ymin;
    a.x[2] = box.xmax; a.y[2] = box.ymax;
    a.x[3] = box.xmin; a.y[3] = box.ymax;
}

// box_area_of_intersection, algorithms/sutherland_hodgman.py:171-187; flag = area > 0 (celltree.py:183)
template <int MAXB>
__global__ void __launch_bounds__(BLOCK) k_box_area(TreeView t, const double *__restrict__ boxes, const int32_t *__restrict__ pi,
                                                    const int32_t *__restrict__ pj, int64_t n, double *__restrict__ area,
                                                    int32_t *__restrict__ flag) {
    int64_t k = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
    if (k >= n) return;
    Poly<4> a;
    box_polygon(load_box(boxes, pi[k]), a);
    Poly<MAXB> b;
    load_polygon<MAXB>(t.elements, t.M, pj[k], t.vertices, b);
    double ar = polygon_polygon_clip_area<4, MAXB>(a, b);
    area[k] = ar;
    flag[k] = ar > 0 ? 1 : 0;
}

// polygons_intersect, algorithms/separating_axis.py:58-75
template <int MAXA, int MAXB>
__global__ void __launch_bounds__(BLOCK) k_sat(TreeView t, const int32_t *__restrict__ qfaces, int qM,
                                               const double2 *__restrict__ qvertices, const int32_t *__restrict__ pi,
                                               const int32_t *__restrict__ pj, int64_t n, int32_t *__restrict__ flag) {
    int64_t k = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
    if (k >= n) return;
    Poly<MAXA> a;
    load_polygon<MAXA>(qfaces, qM, pi[k], qvertices, a);
    Poly<MAXB> b;
    load_polygon<MAXB>(t.elements, t.M, pj[k], t.vertices, b);
    flag[k] = (separating_axes<MAXA, MAXB>(a, b) && separating_axes<MAXB, MAXA>(b, a)) ? 1 : 0;
}

// area_of_intersection, algorithms/sutherland_hodgman.py:151-168; flag = area > 0 (celltree.py:268)
template <int MAXA, int MAXB>
__global__ void __launch_bounds__(BLOCK) k_clip_area(TreeView t, const int32_t *__restrict__ qfaces, int qM,
                                                     const double2 *__restrict__ qvertices, const int32_t *__restrict__ pi,
                                                     const int32_t *__restrict__ pj, int64_t n, double *__restrict__ area,
                                                     int32_t *__restrict__ flag) {
    int64_t k = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
    if (k >= n) return;
    Poly<MAXA> a;
    load_polygon<MAXA>(qfaces, qM, pi[k], qvertices, a);
    Poly<MAXB> b;
    load_polygon<MAXB>(t.elements, t.M, pj[k], t.vertices, b);
    double ar = polygon_polygon_clip_area<MAXA, MAXB>(a, b);
    area[k] = ar;
    flag[k] = ar > 0 ? 1 : 0;
}

// keep the flagged pairs, order preserved (the NumPy boolean masks of celltree.py:183-184, 226, 268-269)
__global__ void __launch_bounds__(256) k_compact(const int32_t *__restrict__ flag, const int64_t *__restrict__ pos, int64_t n,
                                                 const int32_t *__restrict__ in_i, const int32_t *__restrict__ in_j,
                                                 const double *__restrict__ in_p, int32_t *__restrict__ out_i,
                                                 int32_t *__restrict__ out_j, double *__restrict__ out_p) {
    int64_t k = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (k >= n || !flag[k]) return;
    int64_t o = pos[k];
    out_i[o] = in_i[k];
    out_j[o] = in_j[k];
    if (in_p) out_p[o] = in_p[k];
}

int launch_widen(const int32_t *in, int64_t n, int64_t *out, cudaStream_t s);   // build.cu
int launch_narrow(const int64_t *in, int64_t n, int32_t *out, cudaStream_t s);  // build.cu

// exclusive scan of int32 counts into int64 offsets[n + 1]; total returned through the last element
static int scan_counts(const int32_t *counts, int64_t n, int64_t *offsets, int64_t *total, cudaStream_t s) {
    // offsets[0..n) = exclusive sum; offsets[n] = total.  Scan n + 1 items (counts has a zero sentinel at [n]).
    size_t bytes = 0;
    auto in = cub::TransformInputIterator<int64_t, cub::CastOp<int64_t>, const int32_t *>(counts, cub::CastOp<int64_t>());
    CT_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, offsets, n + 1, s));
    Scratch<char> tmp;
    CT_CHECK(tmp.alloc(bytes, s));
    CT_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, bytes, in, offsets, n + 1, s));
    count_launch(2);
    CT_CUDA(cudaMemcpyAsync(total, offsets + n, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    CT_CUDA(cudaStreamSynchronize(s));
    return CT_OK;
}

// Replace (i, j, payload) of `r` by the flagged subset.
static int compact_result(ct_result *r, const int32_t *flag, bool keep_payload, cudaStream_t s) {
    int64_t n = r->size;
    Scratch<int64_t> pos;
    CT_CHECK(pos.alloc(n + 1, s));
    int64_t total = 0;
    // flag has n entries; scan n+1 needs a sentinel: scan n items and add the last flag on the host side
    {
        size_t bytes = 0;
        auto in = cub::TransformInputIterator<int64_t, cub::CastOp<int64_t>, const int32_t *>(flag, cub::CastOp<int64_t>());
        CT_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, pos.p, n, s));
        Scratch<char> tmp;
        CT_CHECK(tmp.alloc(bytes, s));
        if (n > 0) {
            CT_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, bytes, in, pos.p, n, s));
            count_launch(2);
            int64_t last_pos = 0;
            int32_t last_flag = 0;
            CT_CUDA(cudaMemcpyAsync(&last_pos, pos.p + n - 1, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
            CT_CUDA(cudaMemcpyAsync(&last_flag, flag + n - 1, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
            CT_CUDA(cudaStreamSynchronize(s));
            total = last_pos + last_flag;
        }
    }
    int32_t *ni = nullptr, *nj = nullptr;
    double *np_ = nullptr;
    CT_CHECK(dalloc(&ni, total, s));
    CT_CHECK(dalloc(&nj, total, s));
    if (keep_payload) CT_CHECK(dalloc(&np_, total, s));
    if (n > 0) {
        k_compact<<<grid_for(n, 256), 256, 0, s>>>(flag, pos.p, n, r->i, r->j, keep_payload ? r->payload : nullptr, ni, nj, np_);
        CT_LAUNCH_CHECK();
    }
    dfree(r->i, s);
    dfree(r->j, s);
    dfree(r->payload, s);
    r->i = ni;
    r->j = nj;
    r->payload = np_;
    r->width = keep_payload ? 1 : 0;
    r->size = total;
    return CT_OK;
}

// count -> scan -> fill for boxes already on the device; result pairs in r (int32 i, j)
static int locate_boxes_device(const ct_tree *tree, const double *d_boxes, int64_t n, ct_result *r, cudaStream_t s) {
    TreeView v = tree->view();
    Scratch<int32_t> counts;
    Scratch<int64_t> offsets;
    CT_CHECK(counts.alloc(n + 1, s));
    CT_CHECK(offsets.alloc(n + 1, s));
    CT_CUDA(cudaMemsetAsync(counts.p + n, 0, sizeof(int32_t), s));
    int64_t total = 0;
    if (n > 0) {
        k_locate_boxes<false><<<grid_for(n, BLOCK), BLOCK, 0, s>>>(v, d_boxes, n, counts.p, nullptr, nullptr, nullptr);
        CT_LAUNCH_CHECK();
    }
    CT_CHECK(scan_counts(counts.p, n, offsets.p, &total, s));
    CT_CHECK(dalloc(&r->i, total, s));
    CT_CHECK(dalloc(&r->j, total, s));
    r->size = total;
    r->width = 0;
    if (n > 0 && total > 0) {
        k_locate_boxes<true><<<grid_for(n, BLOCK), BLOCK, 0, s>>>(v, d_boxes, n, nullptr, offsets.p, r->i, r->j);
        CT_LAUNCH_CHECK();
    }
    return CT_OK;
}

template <int MAXB>
static int launch_box_area(const ct_tree *tree, const double *d_boxes, ct_result *r, double *area, int32_t *flag, cudaStream_t s) {
    k_box_area<MAXB><<<grid_for(r->size, BLOCK), BLOCK, 0, s>>>(tree->view(), d_boxes, r->i, r->j, r->size, area, flag);
    CT_LAUNCH_CHECK();
    return CT_OK;
}

template <int MAXA, int MAXB>
static int launch_sat(const ct_tree *tree, const int32_t *qf, int qM, const double2 *qv, ct_result *r, int32_t *flag, cudaStream_t s) {
    k_sat<MAXA, MAXB><<<grid_for(r->size, BLOCK), BLOCK, 0, s>>>(tree->view(), qf, qM, qv, r->i, r->j, r->size, flag);
    CT_LAUNCH_CHECK();
    return CT_OK;
}

template <int MAXA, int MAXB>
static int launch_clip(const ct_tree *tree, const int32_t *qf, int qM, const double2 *qv, ct_result *r, double *area, int32_t *flag,
                       cudaStream_t s) {
    k_clip_area<MAXA, MAXB><<<grid_for(r->size, BLOCK), BLOCK, 0, s>>>(tree->view(), qf, qM, qv, r->i, r->j, r->size, area, flag);
    CT_LAUNCH_CHECK();
    return CT_OK;
}

static inline int bound_of(int M) { return M == 3 ? 3 : (M == 4 ? 4 : 32); }

#define CT_DISPATCH_PAIR(FN, MA, MB, ...)                                   \
    do {                                                                    \
        int _a = bound_of(MA), _b = bound_of(MB);                           \
        if (_a == 3 && _b == 3) CT_CHECK((FN<3, 3>(__VA_ARGS__)));          \
        else if (_a == 3 && _b == 4) CT_CHECK((FN<3, 4>(__VA_ARGS__)));     \
        else if (_a == 3) CT_CHECK((FN<3, 32>(__VA_ARGS__)));               \
        else if (_a == 4 && _b == 3) CT_CHECK((FN<4, 3>(__VA_ARGS__)));     \
        else if (_a == 4 && _b == 4) CT_CHECK((FN<4, 4>(__VA_ARGS__)));     \
        else if (_a == 4) CT_CHECK((FN<4, 32>(__VA_ARGS__)));               \
        else if (_b == 3) CT_CHECK((FN<32, 3>(__VA_ARGS__)));               \
        else if (_b == 4) CT_CHECK((FN<32, 4>(__VA_ARGS__)));               \
        else CT_CHECK((FN<32, 32>(__VA_ARGS__)));                           \
    } while (0)

}  // namespace ct

extern "C" int ct_locate_boxes(const ct_tree *tree, const double *boxes, int64_t n, int32_t with_area, int32_t mem, ct_result **out) {
    if (!tree || !out || n < 0 || (n > 0 && !boxes)) {
        set_error("ct_locate_boxes: null argument");
        return CT_ERR_VALUE;
    }
    if (tree->kind != CT_KIND_FACES && with_area) {
        set_error("ct_locate_boxes: areas need a face tree");
        return CT_ERR_VALUE;
    }
    CT_CHECK(check_depth(tree));
    CT_CUDA(cudaSetDevice(tree->device));
    cudaStream_t s = current_stream();
    DevIn<double> d_boxes;
    CT_CHECK(d_boxes.init(boxes, (size_t)n * 4, mem, s));
    ct_result *r = new ct_result();
    int status = locate_boxes_device(tree, d_boxes.p, n, r, s);
    if (status == CT_OK && with_area) {
        auto stage = [&]() -> int {
            Scratch<double> area;
            Scratch<int32_t> flag;
            CT_CHECK(area.alloc(r->size, s));
            CT_CHECK(flag.alloc(r->size, s));
            if (r->size > 0) {
                int b = tree->M == 3 ? 3 : (tree->M == 4 ? 4 : 32);
                if (b == 3) CT_CHECK(launch_box_area<3>(tree, d_boxes.p, r, area.p, flag.p, s));
                else if (b == 4) CT_CHECK(launch_box_area<4>(tree, d_boxes.p, r, area.p, flag.p, s));
                else CT_CHECK(launch_box_area<32>(tree, d_boxes.p, r, area.p, flag.p, s));
            }
            r->payload = area.release();
            r->width = 1;
            return compact_result(r, flag.p, true, s);
        };
        status = stage();
    }
    if (status == CT_OK) {
        cudaError_t e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) {
            set_error(std::string("ct_locate_boxes: ") + cudaGetErrorString(e));
            status = CT_ERR_CUDA;
        }
    }
    if (status != CT_OK) {
        ct_result_free(r);
        return status;
    }
    *out = r;
    return CT_OK;
}

namespace ct {
// shared with build.cu
int launch_counter_clockwise(const double2 *vertices, int32_t *faces, int64_t n_face, int M, cudaStream_t s);
int launch_face_bboxes(const double2 *vertices, const int32_t *faces, int64_t n_face, int M, double *bb, cudaStream_t s);
}  // namespace ct

extern "C" int ct_locate_faces(const ct_tree *tree, const double *vertices, int64_t n_vertex, int64_t *faces, int64_t n_face,
                               int32_t n_max_vert, int32_t with_area, int32_t mem, ct_result **out) {
    if (!tree || !out || n_face < 0 || n_vertex < 0 || (n_face > 0 && (!faces || !vertices))) {
        set_error("ct_locate_faces: null argument");
        return CT_ERR_VALUE;
    }
    if (tree->kind != CT_KIND_FACES) {
        set_error("ct_locate_faces: needs a face tree");
        return CT_ERR_VALUE;
    }
    if (n_max_vert < 3 || n_max_vert > MAX_N_VERTEX) {
        set_error("ct_locate_faces: faces must have 3..32 columns");
        return CT_ERR_VALUE;
    }
    CT_CHECK(check_depth(tree));
    CT_CUDA(cudaSetDevice(tree->device));
    cudaStream_t s = current_stream();
    const int qM = n_max_vert;
    DevIn<double> d_qv;
    CT_CHECK(d_qv.init(vertices, (size_t)n_vertex * 2, mem, s));
    DevIn<int64_t> d_qf64;
    CT_CHECK(d_qf64.init(faces, (size_t)n_face * qM, mem, s));
    Scratch<int32_t> qf;
    Scratch<double> qbb;
    CT_CHECK(qf.alloc((size_t)n_face * qM, s));
    CT_CHECK(qbb.alloc((size_t)n_face * 4, s));
    const double2 *qv = reinterpret_cast<const double2 *>(d_qv.p);
    ct_result *r = new ct_result();
    auto body = [&]() -> int {
        if (n_face > 0) {
            CT_CHECK(launch_narrow(d_qf64.p, n_face * qM, qf.p, s));
            // counter_clockwise on the query faces, in place as the reference does (celltree.py:212)
            CT_CHECK(launch_counter_clockwise(qv, qf.p, n_face, qM, s));
            int64_t *faces_dev = const_cast<int64_t *>(d_qf64.p);
            CT_CHECK(launch_widen(qf.p, n_face * qM, faces_dev, s));
            if (mem == CT_MEM_HOST)
                CT_CUDA(cudaMemcpyAsync(faces, faces_dev, (size_t)n_face * qM * sizeof(int64_t), cudaMemcpyDeviceToHost, s));
            CT_CHECK(launch_face_bboxes(qv, qf.p, n_face, qM, qbb.p, s));
        }
        CT_CHECK(locate_boxes_device(tree, qbb.p, n_face, r, s));
        {
            Scratch<int32_t> flag;
            CT_CHECK(flag.alloc(r->size, s));
            if (r->size > 0) CT_DISPATCH_PAIR(launch_sat, qM, tree->M, tree, qf.p, qM, qv, r, flag.p, s);
            CT_CHECK(compact_result(r, flag.p, false, s));
        }
        if (with_area) {
            Scratch<double> area;
            Scratch<int32_t> flag;
            CT_CHECK(area.alloc(r->size, s));
            CT_CHECK(flag.alloc(r->size, s));
            if (r->size > 0) CT_DISPATCH_PAIR(launch_clip, qM, tree->M, tree, qf.p, qM, qv, r, area.p, flag.p, s);
            r->payload = area.release();
            r->width = 1;
            CT_CHECK(compact_result(r, flag.p, true, s));
        }
        CT_CUDA(cudaStreamSynchronize(s));
        return CT_OK;
    };
    int status = body();
    if (status != CT_OK) {
        ct_result_free(r);
        return status;
    }
    *out = r;
    return CT_OK;
}

namespace ct {
template <int MAXV>
static int run_edges(const ct_tree *tree, const double *d_edges, int64_t n, ct_result *r, cudaStream_t s) {
    TreeView v = tree->view();
    Scratch<int32_t> counts;
    Scratch<int64_t> offsets;
    CT_CHECK(counts.alloc(n + 1, s));
    CT_CHECK(offsets.alloc(n + 1, s));
    CT_CUDA(cudaMemsetAsync(counts.p + n, 0, sizeof(int32_t), s));
    int64_t total = 0;
    if (n > 0) {
        k_locate_edges<MAXV, false><<<grid_for(n, BLOCK), BLOCK, 0, s>>>(v, d_edges, n, counts.p, nullptr, nullptr, nullptr, nullptr);
        CT_LAUNCH_CHECK();
    }
    CT_CHECK(scan_counts(counts.p, n, offsets.p, &total, s));
    CT_CHECK(dalloc(&r->i, total, s));
    CT_CHECK(dalloc(&r->j, total, s));
    CT_CHECK(dalloc(&r->payload, total * 4, s));
    r->size = total;
    r->width = 4;
    if (n > 0 && total > 0) {
        k_locate_edges<MAXV, true><<<grid_for(n, BLOCK), BLOCK, 0, s>>>(v, d_edges, n, nullptr, offsets.p, r->i, r->j, r->payload);
        CT_LAUNCH_CHECK();
        k_sort_edge_ranges<<<grid_for(n, BLOCK), BLOCK, 0, s>>>(d_edges, n, offsets.p, r->j, r->payload);
        CT_LAUNCH_CHECK();
    }
    CT_CUDA(cudaStreamSynchronize(s));
    return CT_OK;
}
}  // namespace ct

extern "C" int ct_intersect_edges(const ct_tree *tree, const double *edges, int64_t n, int32_t mem, ct_result **out) {
    if (!tree || !out || n < 0 || (n > 0 && !edges)) {
        set_error("ct_intersect_edges: null argument");
        return CT_ERR_VALUE;
    }
    CT_CHECK(check_depth(tree));
    CT_CUDA(cudaSetDevice(tree->device));
    cudaStream_t s = current_stream();
    DevIn<double> d_edges;
    CT_CHECK(d_edges.init(edges, (size_t)n * 4, mem, s));
    ct_result *r = new ct_result();
    int status;
    if (tree->kind == CT_KIND_EDGES) status = run_edges<0>(tree, d_edges.p, n, r, s);
    else if (tree->M == 3) status = run_edges<3>(tree, d_edges.p, n, r, s);
    else if (tree->M == 4) status = run_edges<4>(tree, d_edges.p, n, r, s);
    else if (tree->M <= 8) status = run_edges<8>(tree, d_edges.p, n, r, s);
    else status = run_edges<32>(tree, d_edges.p, n, r, s);
    if (status != CT_OK) {
        ct_result_free(r);
        return status;
    }
    *out = r;
    return CT_OK;
}

extern "C" int64_t ct_result_size(const ct_result *r) { return r ? r->size : 0; }
extern "C" int32_t ct_result_payload_width(const ct_result *r) { return r ? r->width : 0; }

extern "C" int ct_result_fetch(const ct_result *r, int64_t *i, int64_t *j, double *payload, int32_t mem) {
    if (!r) {
        set_error("ct_result_fetch: null result");
        return CT_ERR_VALUE;
    }
    cudaStream_t s = current_stream();
    int64_t n = r->size;
    if (n == 0) return CT_OK;
    DevOut<int64_t> oi, oj;
    CT_CHECK(oi.init(i, n, mem, s));
    CT_CHECK(oj.init(j, n, mem, s));
    if (i) {
        CT_CHECK(launch_widen(r->i, n, oi.p, s));
        CT_CHECK(oi.finish(s));
    }
    if (j) {
        CT_CHECK(launch_widen(r->j, n, oj.p, s));
        CT_CHECK(oj.finish(s));
    }
    if (payload && r->width > 0)
        CT_CUDA(cudaMemcpyAsync(payload, r->payload, (size_t)n * r->width * sizeof(double),
                                mem == CT_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, s));
    CT_CUDA(cudaStreamSynchronize(s));
    return CT_OK;
}

extern "C" void ct_result_free(ct_result *r) {
    if (!r) return;
    cudaStream_t s = current_stream();
    dfree(r->i, s);
    dfree(r->j, s);
    dfree(r->payload, s);
    delete r;
}

extern "C" const char *ct_last_error(void) { return g_error.c_str(); }

extern "C" int ct_device_count(int *count) {
    CT_CUDA(cudaGetDeviceCount(count));
    return CT_OK;
}

extern "C" int ct_set_device(int device) {
    CT_CUDA(cudaSetDevice(device));
    return CT_OK;
}

extern "C" int ct_set_stream(void *cuda_stream) {
    g_stream = (cudaStream_t)cuda_stream;
    return CT_OK;
}

extern "C" int64_t ct_launch_count(void) { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }
