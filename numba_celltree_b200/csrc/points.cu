// points.cu -- ct_locate_points: Morton ordering of the queries, the traversal kernel (entry grid, treelet descent,
// point-in-polygon test, fused barycentric weights at the hit; four points per thread), results written in execution
// order and un-permuted through one radix pass; for host buffers a chunked three-stream pipeline around it.
#include "morton.cuh"
#include "traverse.cuh"

namespace ct {

// Queries per thread.  A thread's queries are gathered together up front: their indices (coalesced reads of `perm`),
// then their points as asynchronous 16-byte copies into shared memory that are all in flight at once -- the two
// dependent DRAM round trips (index, then point) are paid once per PER_THREAD queries instead of once per query.
constexpr int PER_THREAD = 4;  // measured on C2, traversal ms per 100 M points: 2 -> 3.97, 4 -> 3.79, 8 -> 3.84

CT_DEV void async_copy_16(void *smem_dst, const void *gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src)
                 : "memory");
}
CT_DEV void async_copy_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }

template <int MAXV, bool WEIGHTS>
CT_DEV void write_weights(const TreeView &t, int found, P2 p, double tolerance, double *__restrict__ w_out) {
    const int M = t.M;
    Poly<MAXV> poly;  // the hit face again (its lines are in L1 from the test a moment ago)
    if (found != -1) load_polygon<MAXV>(t.elements, M, found, t.elem_xy, poly);
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

template <int MAXV, bool WEIGHTS, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB) k_locate_points(TreeView t, const double2 *__restrict__ points, int64_t n,
                                                         double tolerance, int64_t *__restrict__ out,
                                                         double *__restrict__ weights, const uint32_t *__restrict__ perm,
                                                         int32_t *__restrict__ out_in_order) {
    __shared__ double2 s_points[PER_THREAD * BLOCK];
    __shared__ uint32_t s_index[PER_THREAD * BLOCK];
    const int64_t first = (int64_t)blockIdx.x * (PER_THREAD * BLOCK) + threadIdx.x;
    // every thread touches only its own shared-memory entries: no barrier is needed
#pragma unroll
    for (int k = 0; k < PER_THREAD; k++) {
        const int64_t slot = first + (int64_t)k * BLOCK;
        if (slot < n) {
            // perm / points / results are touched once: streaming accesses keep L1 for the tree
            int64_t i = slot;
            if (perm) {
                const uint32_t j = __ldcs(perm + slot);
                s_index[k * BLOCK + threadIdx.x] = j;
                i = j;
            }
            async_copy_16(&s_points[k * BLOCK + threadIdx.x], points + i);
        }
    }
    async_copy_wait();
#pragma unroll 1
    for (int k = 0; k < PER_THREAD; k++) {
        const int64_t slot = first + (int64_t)k * BLOCK;
        if (slot >= n) break;
        const int64_t i = perm ? (int64_t)s_index[k * BLOCK + threadIdx.x] : slot;
        const double2 pt = s_points[k * BLOCK + threadIdx.x];
        const P2 p{pt.x, pt.y};
        const int found = locate_point<MAXV>(t, p, tolerance);
        if (out_in_order) out_in_order[slot] = found;  // coalesced; MortonOrder::scatter_results puts it in place
        else __stcs(out + i, (int64_t)found);
        if constexpr (WEIGHTS) write_weights<MAXV, WEIGHTS>(t, found, p, tolerance, weights + i * (int64_t)t.M);
    }
}

__global__ void __launch_bounds__(BLOCK) k_locate_points_on_edge(TreeView t, const double2 *__restrict__ points, int64_t n,
                                                                 double tolerance, int64_t *__restrict__ out,
                                                                 const uint32_t *__restrict__ perm, int32_t *__restrict__ out_in_order) {
    const int64_t slot = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
    if (slot >= n) return;
    const int64_t i = perm ? (int64_t)__ldcs(perm + slot) : slot;
    double2 pt = __ldcs(points + i);
    const int found = locate_point_on_edge(t, P2{pt.x, pt.y}, tolerance);
    if (out_in_order) out_in_order[slot] = found;
    else __stcs(out + i, (int64_t)found);
}

template <int MAXV>
static int launch_locate_points(const TreeView &v, const double2 *pts, int64_t n, double tol, int64_t *out, double *weights,
                                const uint32_t *perm, int32_t *in_order, cudaStream_t s) {
    int grid = grid_for(n, PER_THREAD * BLOCK);
    // 56 registers (9 blocks of 128 threads per SM) for the 3- and 4-vertex kernels: measured on C2, ms per 100 M points,
    // 64 regs 6.62, 56 regs 6.55, 48 regs 6.57, 40 regs 6.82 -- the cap hardly matters since the descent loop is lean
    if (weights)
        k_locate_points<MAXV, true, 8><<<grid, BLOCK, 0, s>>>(v, pts, n, tol, out, weights, perm, in_order);
    else if (MAXV <= 4)
        k_locate_points<(MAXV <= 4 ? MAXV : 4), false, 9><<<grid, BLOCK, 0, s>>>(v, pts, n, tol, out, weights, perm, in_order);
    else
        k_locate_points<MAXV, false, 8><<<grid, BLOCK, 0, s>>>(v, pts, n, tol, out, weights, perm, in_order);
    CT_LAUNCH_CHECK();
    return CT_OK;
}

static int locate_points_device(const ct_tree *tree, const double2 *pts, int64_t n, double tol, int64_t *out, double *weights,
                                cudaStream_t s, bool profile = false) {
    if (n == 0) return CT_OK;
    TreeView v = tree->view();
    PhaseEvents *ev = profile ? phase_events() : nullptr;
    if (ev) CT_CUDA(cudaEventRecord(ev->start, s));
    MortonOrder order;
    CT_CHECK(order.build<KEY_POINT>(tree, reinterpret_cast<const double *>(pts), n, s));
    const uint32_t *perm = order.perm;
    static int two_phase = -1;
    if (two_phase < 0) {
        const char *e = getenv("CELLTREE_SCATTER");
        two_phase = (e && e[0] == 'd') ? 0 : 1;  // "direct": results are stored straight to out[perm[t]]
    }
    int32_t *in_order = (perm && two_phase) ? reinterpret_cast<int32_t *>(order.spare[0]) : nullptr;
    if (ev) CT_CUDA(cudaEventRecord(ev->ordered, s));
    int status;
    if (tree->kind == CT_KIND_EDGES) {
        k_locate_points_on_edge<<<grid_for(n, BLOCK), BLOCK, 0, s>>>(v, pts, n, tol, out, perm, in_order);
        CT_LAUNCH_CHECK();
        status = CT_OK;
    } else if (tree->M == 3) status = launch_locate_points<3>(v, pts, n, tol, out, weights, perm, in_order, s);
    else if (tree->M == 4) status = launch_locate_points<4>(v, pts, n, tol, out, weights, perm, in_order, s);
    else if (tree->M <= 8) status = launch_locate_points<8>(v, pts, n, tol, out, weights, perm, in_order, s);
    else status = launch_locate_points<32>(v, pts, n, tol, out, weights, perm, in_order, s);
    if (ev) CT_CUDA(cudaEventRecord(ev->done, s));
    if (status == CT_OK && in_order) status = order.scatter_results(n, out, s);
    return status;
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
        return locate_points_device(tree, reinterpret_cast<const double2 *>(points), n, tolerance, out_index, weights, s, true);

    // host buffers: chunked pipeline on two private streams (copy-in / kernel / copy-out overlap)
    // three streams: while one chunk is being copied in, the previous two may still be computing / copying out, so
    // the host-to-device copy engine -- the bottleneck of this path -- never waits for a stream to come free
    static int64_t CHUNK = 0;
    if (CHUNK == 0) {
        const char *e = getenv("CELLTREE_HOST_CHUNK");
        CHUNK = e ? atoll(e) : (1 << 22);
        if (CHUNK < 1024) CHUNK = 1024;
    }
    static int64_t TAIL_MIN = -1;  // smallest chunk of the ramp-down at the end of a batch; 0 = uniform chunks
    if (TAIL_MIN < 0) {
        const char *e = getenv("CELLTREE_HOST_TAIL");
        TAIL_MIN = e ? atoll(e) : (1 << 16);  // measured on C2: 32.76 ms uniform, 32.47 with 2^18, 32.35 with 2^16
        if (TAIL_MIN < 0) TAIL_MIN = 0;
    }
    constexpr int NS = 3;
    const int M = tree->M;
    cudaStream_t st[NS];
    double2 *d_pts[NS] = {};
    int64_t *d_out[NS] = {};
    double *d_w[NS] = {};
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
        int64_t m = 0;
        for (int64_t lo = 0; lo < n; lo += m, k = (k + 1) % NS) {
            // The host-to-device copy engine is the bottleneck and runs without a gap, so the call ends one chunk's
            // kernels + result copy after the last input byte has arrived: the last chunks are halved down to TAIL_MIN
            // so that this drain is short (full-size chunks everywhere else keep the per-chunk launch cost low).
            const int64_t rest = n - lo;
            m = rest < chunk ? rest : chunk;
            if (TAIL_MIN > 0 && n > chunk && rest < 2 * chunk && rest > TAIL_MIN) {
                m = rest / 2 > TAIL_MIN ? rest / 2 : TAIL_MIN;
                if (m > chunk) m = chunk;
            }
            CT_CHECK(upload_from_host(d_pts[k], points + 2 * lo, m * sizeof(double2), st[k]));
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
        pool_stream_synced(st[k]);
        cudaStreamDestroy(st[k]);
    }
    return status;
}

