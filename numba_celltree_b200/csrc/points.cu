// points.cu -- ct_locate_points.  Large batches: the points are binned by a 16-bit Z-order key (binning.cuh), the
// traversal kernel takes tiles of 2048 binned records, orders each tile in shared memory and walks the tree (entry
// grid, treelet descent, point-in-polygon test, fused barycentric weights at the hit); results return through
// per-window queues.  Small batches: one kernel in the caller's order.  For host buffers a chunked three-stream
// pipeline around either.
#include <cub/block/block_radix_sort.cuh>

#include "binning.cuh"
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

// ---- the traversal over binned records -----------------------------------------------------------------------------
// One block = one tile of TILE consecutive records (binning.cuh).  The tile is sorted by the key bits below the bin
// (as many as the tile spans: 8 + log2(bins in the tile)), then thread t walks the tree for the sorted positions
// t, t + 256, ...: a warp's 32 points are neighbours along the Z-order curve.  MAXV == 0: EdgeCellTree2d.
#ifndef CT_EXP2
#define CT_EXP2 0
#endif
#ifndef CT_TILE_MINB
#define CT_TILE_MINB 4  // blocks per SM of the 3- and 4-vertex kernels (64 registers)
#endif
constexpr int TILE_THREADS = 256;
constexpr int TILE_ITEMS = 8;
constexpr int TILE = TILE_THREADS * TILE_ITEMS;

using TileSort = cub::BlockRadixSort<uint16_t, TILE_THREADS, TILE_ITEMS, uint16_t>;
struct TileShared {
    typename TileSort::TempStorage sort;
    uint32_t span, key0;
};

template <int MAXV, bool WEIGHTS, int MINB>
__global__ void __launch_bounds__(TILE_THREADS, MINB)
    k_locate_points_binned(TreeView t, const PointRecord *__restrict__ records, int64_t n, double tolerance,
                           uint2 *__restrict__ pairs, uint32_t *__restrict__ window_cursor, int64_t *__restrict__ out,
                           double *__restrict__ weights) {
    __shared__ TileShared sh;
    const int64_t base = (int64_t)blockIdx.x * TILE;
    const int m = (int)((n - base) < TILE ? (n - base) : TILE);
    const PointRecord *tile = records + base;
    if (threadIdx.x == 0) {
        sh.span = 0;
        // the records are in bin order, so no key of the tile is below its first record's bin
        sh.key0 = (__ldg(&tile[0].key) >> FINE_BITS) << FINE_BITS;
    }
    __syncthreads();
    // Only the keys are read for the sort (the records stay in L2 for the second read below): key relative to the
    // tile's first bin, saturating.  Which thread holds which item does not matter to a sort; `source` says where
    // the item sits in the tile.
    const uint32_t key0 = sh.key0;
    uint16_t keys[TILE_ITEMS], source[TILE_ITEMS];
    uint32_t span = 0;
#pragma unroll
    for (int k = 0; k < TILE_ITEMS; k++) {
        const int j = k * TILE_THREADS + threadIdx.x;
        uint32_t rel = 0xffffu;
        if (j < m) {
            rel = __ldg(&tile[j].key) - key0;
            rel = rel < 0xffffu ? rel : 0xffffu;
            span = rel > span ? rel : span;
        }
        keys[k] = (uint16_t)rel;
        source[k] = (uint16_t)j;
    }
    span = __reduce_max_sync(0xffffffffu, span);
    if ((threadIdx.x & 31) == 0) atomicMax(&sh.span, span);
    __syncthreads();
#if CT_EXP2 == 2
    const int bits = 0;
#else
    const int bits = 32 - __clz(sh.span | 1u);
#endif
    TileSort(sh.sort).SortBlockedToStriped(keys, source, 0, bits);
    // thread t now holds the sorted positions t, t + 256, ...: a warp's 32 points are neighbours on the Z-order curve.
    // The record of the next position is requested while the tree is walked for the current one, and so is the
    // slot in the result queue of the query's window (an atomic whose answer is only needed after the walk).
    uint64_t order_lo = 0, order_hi = 0;  // source[] packed, so that the loop below can stay rolled without a local array
#pragma unroll
    for (int k = 0; k < TILE_ITEMS; k++) {
        if (k < 4) order_lo |= (uint64_t)source[k] << (16 * k);
        else order_hi |= (uint64_t)source[k] << (16 * (k - 4));
    }
    static_assert(TILE_ITEMS == 8, "source[] is packed into two 64-bit words");
    double x, y;
    uint32_t index, key;
    int j = (int)(order_lo & 0xffffu);
    if (j < m) load_record(tile + j, x, y, index, key);
#pragma unroll 1
    for (int k = 0; k < TILE_ITEMS; k++) {
        const bool valid = j < m;
        const P2 p{x, y};
        const uint32_t my_index = index;
        if (k + 1 < TILE_ITEMS) {
            j = (int)(((k + 1 < 4 ? order_lo : order_hi) >> (16 * ((k + 1) & 3))) & 0xffffu);
            if (j < m) load_record(tile + j, x, y, index, key);
        }
        if (!valid) continue;
        uint32_t slot = 0;
#if CT_EXP2 != 1
        if (pairs) slot = atomicAdd(window_cursor + (my_index >> WINDOW_BITS), 1u);
#endif
        int found;
        if constexpr (MAXV == 0) found = locate_point_on_edge(t, p, tolerance);
        else found = locate_point<MAXV>(t, p, tolerance);
        // (index, result) -> the queue of the index's window; small batches: straight to out (L2 merges the stores)
#if CT_EXP2 == 1
        if (pairs) pairs[base + k * TILE_THREADS + threadIdx.x] = make_uint2(my_index, (uint32_t)found);
#else
        if (pairs) pairs[((int64_t)(my_index >> WINDOW_BITS) << WINDOW_BITS) + slot] = make_uint2(my_index, (uint32_t)found);
#endif
        else out[my_index] = (int64_t)found;
        if constexpr (WEIGHTS) write_weights<MAXV, WEIGHTS>(t, found, p, tolerance, weights + (int64_t)my_index * t.M);
    }
}

template <int MAXV>
static int launch_locate_points_binned(const TreeView &v, const PointRecord *records, int64_t n, double tol, uint2 *pairs,
                                       uint32_t *window_cursor, int64_t *out, double *weights, cudaStream_t s) {
    const int grid = grid_for(n, TILE);
    auto launch = [&](auto kernel) -> int {
        kernel<<<grid, TILE_THREADS, 0, s>>>(v, records, n, tol, pairs, window_cursor, out, weights);
        return CT_OK;
    };
    if (weights) {
        if constexpr (MAXV > 0) CT_CHECK(launch(k_locate_points_binned<MAXV, true, 2>));
    } else if (MAXV <= 4)
        CT_CHECK(launch(k_locate_points_binned<MAXV, false, CT_TILE_MINB>));
    else
        CT_CHECK(launch(k_locate_points_binned<MAXV, false, 2>));
    CT_LAUNCH_CHECK();
    return CT_OK;
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

// batches up to this many queries write their results straight to out[index]: the 8-byte stores of a batch whose
// result array (8 n bytes) stays in L2 merge there; larger batches go through the window queues
static int64_t direct_out_limit() {
    static int64_t limit = -1;
    if (limit < 0) {
        const char *e = getenv("CELLTREE_DIRECT_OUT");
        limit = e ? atoll(e) : ((int64_t)1 << 22);
    }
    return limit;
}

static int locate_points_device(const ct_tree *tree, const double2 *pts, int64_t n, double tol, int64_t *out, double *weights,
                                cudaStream_t s, bool profile = false) {
    if (n == 0) return CT_OK;
    TreeView v = tree->view();
    PhaseEvents *ev = profile ? phase_events() : nullptr;
    if (ev) CT_CUDA(cudaEventRecord(ev->start, s));
    const bool binned = sort_bits_for(tree, n) > 0;
    if (!binned) {
        // the caller's order: small batches, small trees
        if (ev) CT_CUDA(cudaEventRecord(ev->ordered, s));
        int status = CT_OK;
        if (tree->kind == CT_KIND_EDGES) {
            k_locate_points_on_edge<<<grid_for(n, BLOCK), BLOCK, 0, s>>>(v, pts, n, tol, out, nullptr, nullptr);
            CT_LAUNCH_CHECK();
        } else if (tree->M == 3) status = launch_locate_points<3>(v, pts, n, tol, out, weights, nullptr, nullptr, s);
        else if (tree->M == 4) status = launch_locate_points<4>(v, pts, n, tol, out, weights, nullptr, nullptr, s);
        else if (tree->M <= 8) status = launch_locate_points<8>(v, pts, n, tol, out, weights, nullptr, nullptr, s);
        else status = launch_locate_points<32>(v, pts, n, tol, out, weights, nullptr, nullptr, s);
        if (ev) CT_CUDA(cudaEventRecord(ev->done, s));
        return status;
    }
    PointBins bins;
    CT_CHECK(bins.build(tree, pts, n, s));
    Scratch<uint2> pairs;
    const bool queued = n > direct_out_limit();
    if (queued) CT_CHECK(pairs.alloc(n, s));
    if (ev) CT_CUDA(cudaEventRecord(ev->ordered, s));
    int status;
    uint32_t *wc = bins.window_cursor();
    if (tree->kind == CT_KIND_EDGES) status = launch_locate_points_binned<0>(v, bins.records.p, n, tol, pairs.p, wc, out, nullptr, s);
    else if (tree->M == 3) status = launch_locate_points_binned<3>(v, bins.records.p, n, tol, pairs.p, wc, out, weights, s);
    else if (tree->M == 4) status = launch_locate_points_binned<4>(v, bins.records.p, n, tol, pairs.p, wc, out, weights, s);
    else if (tree->M <= 8) status = launch_locate_points_binned<8>(v, bins.records.p, n, tol, pairs.p, wc, out, weights, s);
    else status = launch_locate_points_binned<32>(v, bins.records.p, n, tol, pairs.p, wc, out, weights, s);
    if (ev) CT_CUDA(cudaEventRecord(ev->done, s));
    if (status == CT_OK && queued) {
        CT_CUDA(cudaFuncSetAttribute(k_windows_to_out, cudaFuncAttributeMaxDynamicSharedMemorySize, WINDOW * (int)sizeof(int32_t)));
        k_windows_to_out<<<(unsigned)bins.n_windows, WINDOW_THREADS, WINDOW * sizeof(int32_t), s>>>(pairs.p, n, out);
        CT_LAUNCH_CHECK();
    }
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
    CT_ON_DEVICE(tree->device);
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


// Diagnostics: device time of the binning alone (count + offsets + scatter) over `repeats` runs, points on the device.
extern "C" int ct_profile_binning(const ct_tree *tree, const double *points, int64_t n, int32_t repeats, double *ms_per_run) {
    if (!tree || !points || n <= 0 || repeats <= 0 || !ms_per_run) {
        set_error("ct_profile_binning: bad argument");
        return CT_ERR_VALUE;
    }
    CT_ON_DEVICE(tree->device);
    cudaStream_t s = current_stream();
    cudaEvent_t e0, e1;
    CT_CUDA(cudaEventCreate(&e0));
    CT_CUDA(cudaEventCreate(&e1));
    {
        PointBins warm;
        CT_CHECK(warm.build(tree, reinterpret_cast<const double2 *>(points), n, s));
    }
    CT_CUDA(cudaEventRecord(e0, s));
    for (int r = 0; r < repeats; r++) {
        PointBins bins;
        CT_CHECK(bins.build(tree, reinterpret_cast<const double2 *>(points), n, s));
    }
    CT_CUDA(cudaEventRecord(e1, s));
    CT_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    CT_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    *ms_per_run = ms / repeats;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return CT_OK;
}
