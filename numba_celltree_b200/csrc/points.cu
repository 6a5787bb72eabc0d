// points.cu -- ct_locate_points: Morton ordering of the queries, one-thread-per-point traversal with the
// point-in-polygon test and fused barycentric weights at the hit.
#include <cub/device/device_radix_sort.cuh>

#include "traverse.cuh"

namespace ct {

// ---- Morton ordering of the queries -----------------------------------------------------------------------
// Queries are independent, so the order in which threads pick them up is an execution detail: thread t handles
// query perm[t] and writes result slot perm[t].  Sorting the queries along a Z-order curve over the tree's
// bounding box makes the 32 lanes of a warp walk (almost) the same root-to-leaf path, so node / face / vertex
// loads collapse to a few sectors per warp and the lower tree levels are served by L1/L2 instead of HBM.
CT_DEV uint32_t spread16(uint32_t v) {  // 16 bits -> every other bit of 32
    v &= 0xffffu;
    v = (v | (v << 8)) & 0x00ff00ffu;
    v = (v | (v << 4)) & 0x0f0f0f0fu;
    v = (v | (v << 2)) & 0x33333333u;
    v = (v | (v << 1)) & 0x55555555u;
    return v;
}

__global__ void __launch_bounds__(256) k_morton_keys(const double2 *__restrict__ points, int64_t n, double xmin, double ymin,
                                                     double sx, double sy, int shift, uint32_t *__restrict__ keys,
                                                     uint32_t *__restrict__ idx) {
    int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    double2 p = __ldg(points + i);
    double fx = (p.x - xmin) * sx, fy = (p.y - ymin) * sy;  // [0, 65536) inside the tree's bounding box
    fx = fx >= 0.0 ? fx : 0.0;  // also catches NaN
    fy = fy >= 0.0 ? fy : 0.0;
    uint32_t ix = fx < 65535.0 ? (uint32_t)fx : 65535u;
    uint32_t iy = fy < 65535.0 ? (uint32_t)fy : 65535u;
    keys[i] = (spread16(ix) | (spread16(iy) << 1)) >> shift;
    idx[i] = (uint32_t)i;
}

template <int MAXV, bool WEIGHTS, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB) k_locate_points(TreeView t, const double2 *__restrict__ points, int64_t n,
                                                         double tolerance, int64_t *__restrict__ out,
                                                         double *__restrict__ weights, const uint32_t *__restrict__ perm) {
    int64_t i = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
    if (i >= n) return;
    // perm / points / results are touched once: streaming accesses keep L1 for the tree
    if (perm) i = __ldcs(perm + i);
    double2 pt = __ldcs(points + i);
    P2 p{pt.x, pt.y};
    Poly<MAXV> poly;
    int found = locate_point<MAXV>(t, p, tolerance, poly);
    __stcs(out + i, (int64_t)found);
    if constexpr (WEIGHTS) {
        const int M = t.M;
        double *w_out = weights + i * (int64_t)M;
        if constexpr (MAXV == 3) {
            // barycentric_triangle_weights, algorithms/barycentric_triangle.py:46-64
            double u = 0.0, v = 0.0, w = 0.0;
            if (found != -1)
                triangle_weights(P2{poly.x[0], poly.y[0]}, P2{poly.x[1], poly.y[1]}, P2{poly.x[2], poly.y[2]}, p, u, v, w);
            __stcs(w_out, u);
            __stcs(w_out + 1, v);
            __stcs(w_out + 2, w);
        } else {
            // barycentric_wachspress_weights, algorithms/barycentric_wachspress.py:88-107
            double w[MAXV];
#pragma unroll
            for (int k = 0; k < MAXV; k++) w[k] = 0.0;
            if (found != -1) wachspress_weights<MAXV>(poly, p, tolerance, w);
            if constexpr (MAXV == 4) {
                if (M == 4) {
                    double2 *o = reinterpret_cast<double2 *>(w_out);
                    __stcs(o, make_double2(w[0], w[1]));
                    __stcs(o + 1, make_double2(w[2], w[3]));
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
                                                                 double tolerance, int64_t *__restrict__ out,
                                                                 const uint32_t *__restrict__ perm) {
    int64_t i = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
    if (i >= n) return;
    if (perm) i = __ldcs(perm + i);
    double2 pt = __ldcs(points + i);
    __stcs(out + i, (int64_t)locate_point_on_edge(t, P2{pt.x, pt.y}, tolerance));
}

template <int MAXV>
static int launch_locate_points(const TreeView &v, const double2 *pts, int64_t n, double tol, int64_t *out, double *weights,
                                const uint32_t *perm, cudaStream_t s) {
    int grid = grid_for(n, BLOCK);
    // 56 registers (9 blocks of 128 threads per SM) instead of 64: the traversal is latency-bound and one more
    // resident block per SM is worth the handful of spilled values; tighter caps spill into the hot loop and
    // lose (measured on C2, ms per 100 M points: 64 regs 12.7, 56 regs 11.9, 48 regs 14.7, 40 regs 17.3).
    static int minb = -1;
    if (minb < 0) {
        const char *e = getenv("CELLTREE_POINTS_MINB");
        minb = e ? atoi(e) : 9;
    }
    if (weights)
        k_locate_points<MAXV, true, 8><<<grid, BLOCK, 0, s>>>(v, pts, n, tol, out, weights, perm);
    else if (MAXV <= 4 && minb == 9)
        k_locate_points<(MAXV <= 4 ? MAXV : 4), false, 9><<<grid, BLOCK, 0, s>>>(v, pts, n, tol, out, weights, perm);
    else
        k_locate_points<MAXV, false, 8><<<grid, BLOCK, 0, s>>>(v, pts, n, tol, out, weights, perm);
    CT_LAUNCH_CHECK();
    return CT_OK;
}

// Number of Morton key bits to sort the queries by; 0 = keep the caller's order.
// Sorting pays when the tree is much larger than L2 and there are enough queries to amortise the passes.
static int sort_bits_for(const ct_tree *tree, int64_t n) {
    const int forced = sort_bits_override();
    if (n >= (1LL << 31)) return 0;
    if (forced >= 0) return forced > 32 ? 32 : forced;
    const double tree_bytes = 32.0 * (double)tree->n_nodes + (double)tree->n_elem * (4.0 + 4.0 * tree->M) + 16.0 * (double)tree->n_vertex;
    if (n < (1 << 20) || tree_bytes < 48e6) return 0;
    int bits = 8;  // about one key per query, whole 8-bit radix passes, at most three of them
    while (bits < 24 && (1LL << bits) < n) bits += 8;
    return bits;
}

// order[] = indices of the points sorted by Morton key (radix sort of (key, index) pairs)
static int morton_order(const ct_tree *tree, const double2 *pts, int64_t n, int bits, Scratch<uint32_t> &keys_a,
                        Scratch<uint32_t> &keys_b, Scratch<uint32_t> &idx_a, Scratch<uint32_t> &idx_b, const uint32_t **order,
                        cudaStream_t s) {
    CT_CHECK(keys_a.alloc(n, s));
    CT_CHECK(keys_b.alloc(n, s));
    CT_CHECK(idx_a.alloc(n, s));
    CT_CHECK(idx_b.alloc(n, s));
    double wx = tree->bbox[1] - tree->bbox[0], wy = tree->bbox[3] - tree->bbox[2];
    double sx = wx > 0 ? 65536.0 / wx : 0.0, sy = wy > 0 ? 65536.0 / wy : 0.0;
    k_morton_keys<<<grid_for(n, 256), 256, 0, s>>>(pts, n, tree->bbox[0], tree->bbox[2], sx, sy, 32 - bits, keys_a.p, idx_a.p);
    CT_LAUNCH_CHECK();
    cub::DoubleBuffer<uint32_t> d_keys(keys_a.p, keys_b.p);
    cub::DoubleBuffer<uint32_t> d_vals(idx_a.p, idx_b.p);
    size_t bytes = 0;
    CT_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, d_keys, d_vals, n, 0, bits, s));
    Scratch<char> tmp;
    CT_CHECK(tmp.alloc(bytes, s));
    CT_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, bytes, d_keys, d_vals, n, 0, bits, s));
    count_launch(1 + (bits + 7) / 8);
    *order = d_vals.Current();
    return CT_OK;
}

static int locate_points_device(const ct_tree *tree, const double2 *pts, int64_t n, double tol, int64_t *out, double *weights,
                                cudaStream_t s) {
    if (n == 0) return CT_OK;
    TreeView v = tree->view();
    Scratch<uint32_t> keys_a, keys_b, idx_a, idx_b;
    const uint32_t *perm = nullptr;
    int bits = sort_bits_for(tree, n);
    if (bits > 0) CT_CHECK(morton_order(tree, pts, n, bits, keys_a, keys_b, idx_a, idx_b, &perm, s));
    if (tree->kind == CT_KIND_EDGES) {
        k_locate_points_on_edge<<<grid_for(n, BLOCK), BLOCK, 0, s>>>(v, pts, n, tol, out, perm);
        CT_LAUNCH_CHECK();
        return CT_OK;
    }
    if (tree->M == 3) return launch_locate_points<3>(v, pts, n, tol, out, weights, perm, s);
    if (tree->M == 4) return launch_locate_points<4>(v, pts, n, tol, out, weights, perm, s);
    if (tree->M <= 8) return launch_locate_points<8>(v, pts, n, tol, out, weights, perm, s);
    return launch_locate_points<32>(v, pts, n, tol, out, weights, perm, s);
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

