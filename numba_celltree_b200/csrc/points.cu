// points.cu -- ct_locate_points.  Large batches: every point is appended to the slab of its Z-order bin (binning.cuh), the
// traversal kernel takes one bin per block, orders its records in shared memory and walks the tree (entry grid, treelet
// descent, point-in-polygon test, fused barycentric weights at the hit); results return through per-window queues.
// Small batches and trees deeper than the per-thread stack: one kernel in the caller's order.  For host buffers a chunked
// three-stream pipeline around either.
#include <cub/block/block_radix_sort.cuh>

#include "binning.cuh"
#include "tile_order.cuh"
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

// set by the weight kernels where the reference would raise ZeroDivisionError (geometry.cuh); read and cleared by the host
__device__ int g_zero_division;

template <int MAXV, bool WEIGHTS>
CT_DEV void write_weights(const TreeView &t, int found, P2 p, double tolerance, double *__restrict__ w_out) {
    const int M = t.M;
    Poly<MAXV> poly;  // the hit face again (its lines are in L1 from the test a moment ago)
    if (found != -1) load_tree_polygon<MAXV>(t, found, poly);
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
        if (found != -1) wachspress_weights<MAXV>(poly, p, tolerance, w, &g_zero_division);
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

template <int MAXV, bool WEIGHTS, int MINB, bool DEEP>
__global__ void __launch_bounds__(BLOCK, MINB) k_locate_points(TreeView t, const double2 *__restrict__ points, int64_t n,
                                                         double tolerance, int64_t *__restrict__ out,
                                                         double *__restrict__ weights, const uint32_t *__restrict__ perm,
                                                         int32_t *__restrict__ out_in_order, uint2 *__restrict__ pairs = nullptr,
                                                         uint32_t *__restrict__ window_cursor = nullptr) {
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
        uint32_t queue_slot = 0;
        if (pairs) queue_slot = atomicAdd(window_cursor + ((uint32_t)i >> WINDOW_BITS), 1u);
        const int found = locate_point<MAXV, NoProbe, DEEP>(t, p, tolerance);
        if (pairs) pairs[((int64_t)((uint32_t)i >> WINDOW_BITS) << WINDOW_BITS) + queue_slot] = make_uint2((uint32_t)i, (uint32_t)found);
        else if (out_in_order) out_in_order[slot] = found;  // coalesced; MortonOrder::scatter_results puts it in place
        else __stcs(out + i, (int64_t)found);
        if constexpr (WEIGHTS) write_weights<MAXV, WEIGHTS>(t, found, p, tolerance, weights + i * (int64_t)t.M);
    }
}

template <bool DEEP>
__global__ void __launch_bounds__(BLOCK) k_locate_points_on_edge(TreeView t, const double2 *__restrict__ points, int64_t n,
                                                                 double tolerance, int64_t *__restrict__ out,
                                                                 const uint32_t *__restrict__ perm, int32_t *__restrict__ out_in_order) {
    const int64_t slot = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
    if (slot >= n) return;
    const int64_t i = perm ? (int64_t)__ldcs(perm + slot) : slot;
    double2 pt = __ldcs(points + i);
    const int found = locate_point_on_edge<DEEP>(t, P2{pt.x, pt.y}, tolerance);
    if (out_in_order) out_in_order[slot] = found;
    else __stcs(out + i, (int64_t)found);
}

// ---- the traversal over binned records -----------------------------------------------------------------------------
// One block = one tile of TILE consecutive records (binning.cuh).  The tile is sorted by the key bits below the bin
// (as many as the tile spans: 8 + log2(bins in the tile)), then thread t walks the tree for the sorted positions
// t, t + 256, ...: a warp's 32 points are neighbours along the Z-order curve.  MAXV == 0: EdgeCellTree2d.
// Measurement switch (build_ext --variant), off by default: CT_EXP2 = 1: the tile kernel without the atomics of the result
// queues (pairs written in execution order: results unusable), 2: without the tile sort
#ifndef CT_EXP2
#define CT_EXP2 0
#endif
#ifndef CT_WEIGHTS_MINB
#define CT_WEIGHTS_MINB 4  // the same with fused barycentric weights (C2, 100 M points: 2 -> 11.9 ms, 3 -> 11.1, 4 -> 10.6)
#endif
#ifndef CT_PREFETCH_RECORD
#define CT_PREFETCH_RECORD 0  // 1: the next record is asked into L1 by a prefetch instead of being loaded one iteration ahead
                              // (five live values across the walk): fewer spills in the FILTER kernels, but 3.44 against 3.22 ms
#endif
#ifndef CT_TILE_MINB
#define CT_TILE_MINB 4  // blocks per SM of the 3- and 4-vertex kernels (64 registers)
#endif
constexpr int TILE_THREADS = 256;
constexpr int TILE_ITEMS = 8;
constexpr int TILE = TILE_THREADS * TILE_ITEMS;

// The order of a tile's queries: a counting sort (tile_order.cuh) when the relative keys have few bits -- every tile of the
// slab bins, 8 to 10 bits -- and the library's block radix sort for the wide keys of the other execution orders.
using TileSort = cub::BlockRadixSort<uint16_t, TILE_THREADS, TILE_ITEMS, uint16_t>;
using TileCount = TileOrder<TILE_THREADS, TILE_ITEMS>;
static_assert(SlabPlan::MAX_SORT_BITS <= ORDER_MAX_BITS, "a slab's relative keys fit the counting sort");
struct TileShared {
    union {
        typename TileSort::TempStorage sort;
        typename TileCount::Storage count;
    };
    uint32_t lo, hi;
};

// Where a tile's queries come from: binned records (binning.cuh), or a permutation (sorted by the 16-bit bin) over the
// caller's point array, gathered here.
struct TileInput {
    const PointRecord *records;  // binned records, or nullptr
    const uint32_t *perm;        // else: query indices in bin order ...
    const double2 *points;       // ... into the caller's points
    BinGrid grid;
    const uint32_t *slab_fill;   // slab bins (binning.cuh): tile b is bin b, records [b * SLAB, b * SLAB + min(fill[b], SLAB)) ...
    SlabPlan plan;               // ... sorted by the key relative to the bin's first key
    bool crowded_leaves;         // host side only: the tree's leaves hold more than two cells (picks the kernel variant)
};

CT_DEV double2 load_point_once(const double2 *p) {
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}

template <int MAXV, bool WEIGHTS, int MINB, bool GATHER, bool FILTER>
__global__ void __launch_bounds__(TILE_THREADS, MINB)
    k_locate_points_binned(TreeView t, TileInput in, int64_t n, double tolerance, uint2 *__restrict__ pairs,
                           uint32_t *__restrict__ window_cursor, int64_t *__restrict__ out, double *__restrict__ weights) {
    __shared__ TileShared sh;
    const bool slabs = !GATHER && in.slab_fill != nullptr;
    const int64_t base = (int64_t)blockIdx.x * TILE;
    int m;
    if (slabs) {
        const uint32_t fill = __ldg(in.slab_fill + blockIdx.x);
        m = fill < (uint32_t)SLAB ? (int)fill : SLAB;
        if (m == 0) return;
    } else {
        m = (int)((n - base) < TILE ? (n - base) : TILE);
    }
    static_assert(SLAB == TILE, "a slab is one tile");
    const PointRecord *tile = in.records + base;
    const uint32_t *tile_perm = in.perm + base;
    if (threadIdx.x == 0) {
        sh.lo = 0xffffffffu;
        sh.hi = 0;
    }
    __syncthreads();
    // Only the keys are needed for the sort (the records / points stay in L2 for the second read below).  Which thread
    // holds which item does not matter to a sort; `source` says where the item sits in the tile.
    uint32_t key24[TILE_ITEMS];
    uint32_t lo = 0xffffffffu, hi = 0;
    if constexpr (GATHER) {
        uint32_t index[TILE_ITEMS];
#pragma unroll
        for (int k = 0; k < TILE_ITEMS; k++) {
            const int j = k * TILE_THREADS + threadIdx.x;
            index[k] = j < m ? __ldg(tile_perm + j) : 0u;
        }
        double2 pt[TILE_ITEMS];
#pragma unroll
        for (int k = 0; k < TILE_ITEMS; k++) {
            const int j = k * TILE_THREADS + threadIdx.x;
            if (j < m) pt[k] = load_point_once(in.points + index[k]);
        }
#pragma unroll
        for (int k = 0; k < TILE_ITEMS; k++) {
            const int j = k * TILE_THREADS + threadIdx.x;
            key24[k] = j < m ? point_key24(in.grid, pt[k].x, pt[k].y) : 0xffffffffu;
        }
    } else {
#pragma unroll
        for (int k = 0; k < TILE_ITEMS; k++) {
            const int j = k * TILE_THREADS + threadIdx.x;
            key24[k] = j < m ? __ldg(&tile[j].key) : 0xffffffffu;
        }
    }
    uint32_t key0, span;
    if (slabs) {
        // one bin: keys relative to the bin's first key, cut down to the plan's sort bits
        span = (1u << in.plan.sort_bits) - 1u;
        key0 = 0;
        const uint32_t first_key = slab_first_key(blockIdx.x, in.plan.bins);
#pragma unroll
        for (int k = 0; k < TILE_ITEMS; k++) {
            const uint32_t rel = (key24[k] - first_key) >> in.plan.shift;
            key24[k] = rel < span ? rel : span;
        }
    } else {
#pragma unroll
        for (int k = 0; k < TILE_ITEMS; k++) {
            const int j = k * TILE_THREADS + threadIdx.x;
            if (j < m) {
                lo = key24[k] < lo ? key24[k] : lo;
                hi = key24[k] > hi ? key24[k] : hi;
            }
        }
        lo = __reduce_min_sync(0xffffffffu, lo);
        hi = __reduce_max_sync(0xffffffffu, hi);
        if ((threadIdx.x & 31) == 0) {
            atomicMin(&sh.lo, lo);
            atomicMax(&sh.hi, hi);
        }
        __syncthreads();
        // key relative to the tile's first bin, saturating (a tile that spans more than 256 bins is too sparse for the
        // order of its far end to matter)
        key0 = (sh.lo >> FINE_BITS) << FINE_BITS;
        span = sh.hi - key0;
        span = span < 0xffffu ? span : 0xffffu;
    }
#if CT_EXP2 == 2
    const int bits = 0;
#else
    const int bits = 32 - __clz(span | 1u);
#endif
    // Afterwards sh.count.order[0 .. m) lists the tile's places by key, and thread t takes the sorted positions t, t + 256,
    // ...: a warp's 32 points are neighbours on the Z-order curve.
    if (bits <= ORDER_MAX_BITS) {
#pragma unroll
        for (int k = 0; k < TILE_ITEMS; k++) {
            const uint32_t rel = key24[k] - key0;
            key24[k] = rel < span ? rel : span;
        }
        TileCount::sort(sh.count, key24, m, bits);
    } else {
        // wide keys (a sparse tile of the counted bins or of the sorted permutation): valid keys saturate at 0xfffe and the
        // empty places of the last, partial tile carry 0xffff -- compared over all 16 bits there, so that they sort strictly
        // last and the first m sorted positions are exactly the tile's points
        uint16_t keys[TILE_ITEMS], source[TILE_ITEMS];
#pragma unroll
        for (int k = 0; k < TILE_ITEMS; k++) {
            const int j = k * TILE_THREADS + threadIdx.x;
            uint32_t rel = 0xffffu;
            if (j < m) {
                rel = key24[k] - key0;
                rel = rel < 0xfffeu ? rel : 0xfffeu;
            }
            keys[k] = (uint16_t)rel;
            source[k] = (uint16_t)j;
        }
        TileSort(sh.sort).SortBlockedToStriped(keys, source, 0, m < TILE ? 16 : bits);
        __syncthreads();
#pragma unroll
        for (int k = 0; k < TILE_ITEMS; k++) sh.count.order[k * TILE_THREADS + threadIdx.x] = source[k];  // places >= m sort last
        __syncthreads();
    }
    // The record of the next position is requested while the tree is walked for the current one, and so is the slot in the
    // result queue of the query's window (an atomic whose answer is only needed after the walk).
    double x = 0.0, y = 0.0;
    uint32_t index = 0;
    auto fetch = [&](int j) {
        if constexpr (GATHER) {
            index = __ldg(tile_perm + j);
            const double2 pt = load_point_once(in.points + index);
            x = pt.x, y = pt.y;
        } else {
            uint32_t key;
            load_record(tile + j, x, y, index, key);
        }
    };
    auto ask = [&](int position) {
        if constexpr (!GATHER) {
            if (position < m) asm volatile("prefetch.global.L1 [%0];" ::"l"(tile + sh.count.order[position]));
        }
    };
#if CT_PREFETCH_RECORD
    ask(threadIdx.x);
#else
    if ((int)threadIdx.x < m) fetch(sh.count.order[threadIdx.x]);
#endif
#pragma unroll 1
    for (int position = threadIdx.x; position < m;) {
#if CT_EXP2 == 1
        const int k = position / TILE_THREADS;
#endif
#if CT_PREFETCH_RECORD
        fetch(sh.count.order[position]);
        position += TILE_THREADS;
        ask(position);
        const P2 p{x, y};
        const uint32_t my_index = index;
#else
        const P2 p{x, y};
        const uint32_t my_index = index;
        position += TILE_THREADS;
        if (position < m) fetch(sh.count.order[position]);
#endif
        uint32_t slot = 0;
#if CT_EXP2 != 1
        if (pairs) slot = atomicAdd(window_cursor + (my_index >> WINDOW_BITS), 1u);
#endif
        int found;
        if constexpr (MAXV == 0) found = locate_point_on_edge(t, p, tolerance);
        else found = locate_point<MAXV, NoProbe, false, FILTER>(t, p, tolerance);
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

// The points that found their bin's slab full (binning.cuh): walked in arrival order by a grid that covers the machine
// once, however many there are (the count stays on the device).
template <int MAXV, bool WEIGHTS>
__global__ void __launch_bounds__(BLOCK) k_locate_points_overflow(TreeView t, const double2 *__restrict__ points, const uint32_t *__restrict__ list,
                                                                  const uint32_t *__restrict__ count, double tolerance,
                                                                  uint2 *__restrict__ pairs, uint32_t *__restrict__ window_cursor,
                                                                  int64_t *__restrict__ out, double *__restrict__ weights) {
    const uint32_t n = __ldg(count);
    for (uint32_t k = blockIdx.x * BLOCK + threadIdx.x; k < n; k += gridDim.x * BLOCK) {
        const uint32_t index = __ldg(list + k);
        const double2 pt = load_point_once(points + index);
        const P2 p{pt.x, pt.y};
        int found;
        if constexpr (MAXV == 0) found = locate_point_on_edge(t, p, tolerance);
        else found = locate_point<MAXV>(t, p, tolerance);
        if (pairs) {
            const uint32_t slot = atomicAdd(window_cursor + (index >> WINDOW_BITS), 1u);
            pairs[((int64_t)(index >> WINDOW_BITS) << WINDOW_BITS) + slot] = make_uint2(index, (uint32_t)found);
        } else {
            out[index] = (int64_t)found;
        }
        if constexpr (WEIGHTS) write_weights<MAXV, WEIGHTS>(t, found, p, tolerance, weights + (int64_t)index * t.M);
    }
}

template <int MAXV>
static int launch_locate_points_overflow(const TreeView &v, const double2 *pts, const uint32_t *list, const uint32_t *count, double tol,
                                         uint2 *pairs, uint32_t *window_cursor, int64_t *out, double *weights, cudaStream_t s) {
    int device = 0, sms = 0;
    CT_CUDA(cudaGetDevice(&device));
    CT_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    const int grid = sms * 8;
    if (weights) {
        if constexpr (MAXV > 0) k_locate_points_overflow<MAXV, true><<<grid, BLOCK, 0, s>>>(v, pts, list, count, tol, pairs, window_cursor, out, weights);
    } else {
        k_locate_points_overflow<MAXV, false><<<grid, BLOCK, 0, s>>>(v, pts, list, count, tol, pairs, window_cursor, out, weights);
    }
    CT_LAUNCH_CHECK();
    return CT_OK;
}

template <int MAXV>
static int launch_locate_points_binned(const TreeView &v, const TileInput &in, int64_t n, double tol, uint2 *pairs,
                                       uint32_t *window_cursor, int64_t *out, double *weights, cudaStream_t s) {
    const int grid = in.slab_fill ? (int)in.plan.bins : grid_for(n, TILE);
    auto launch = [&](auto binned, auto gathered) -> int {
        if (in.records) binned<<<grid, TILE_THREADS, 0, s>>>(v, in, n, tol, pairs, window_cursor, out, weights);
        else gathered<<<grid, TILE_THREADS, 0, s>>>(v, in, n, tol, pairs, window_cursor, out, weights);
        return CT_OK;
    };
    // FILTER (traverse.cuh: the bounding test before the point-in-polygon test): always for the fused weights (C2: 9.85 ->
    // 9.54 ms) and the generic polygons; for the plain 3- and 4-vertex kernels only when a leaf holds more than two cells --
    // with the default two it saves 11 % of the instructions and no time, and costs the kernel its spill-free 64 registers
    if (weights) {
        if constexpr (MAXV > 0) CT_CHECK(launch(k_locate_points_binned<MAXV, true, (MAXV <= 4 ? CT_WEIGHTS_MINB : 2), false, true>, k_locate_points_binned<MAXV, true, (MAXV <= 4 ? CT_WEIGHTS_MINB : 2), true, true>));
    } else if (MAXV <= 4 && in.crowded_leaves)
        CT_CHECK(launch(k_locate_points_binned<MAXV, false, CT_TILE_MINB, false, true>, k_locate_points_binned<MAXV, false, CT_TILE_MINB, true, true>));
    else if (MAXV <= 4)
        CT_CHECK(launch(k_locate_points_binned<MAXV, false, CT_TILE_MINB, false, false>, k_locate_points_binned<MAXV, false, CT_TILE_MINB, true, false>));
    else
        CT_CHECK(launch(k_locate_points_binned<MAXV, false, 2, false, true>, k_locate_points_binned<MAXV, false, 2, true, true>));
    CT_LAUNCH_CHECK();
    return CT_OK;
}

template <int MAXV>
static int launch_locate_points(const TreeView &v, const double2 *pts, int64_t n, double tol, int64_t *out, double *weights,
                                const uint32_t *perm, int32_t *in_order, cudaStream_t s, uint2 *pairs = nullptr,
                                uint32_t *window_cursor = nullptr) {
    int grid = grid_for(n, PER_THREAD * BLOCK);
    // 56 registers (9 blocks of 128 threads per SM) for the 3- and 4-vertex kernels: measured on C2, ms per 100 M points,
    // 64 regs 6.62, 56 regs 6.55, 48 regs 6.57, 40 regs 6.82 -- the cap hardly matters since the descent loop is lean
    if (v.deep.slab) {  // a tree deeper than the per-thread stack
        if (weights) k_locate_points<MAXV, true, 4, true><<<grid, BLOCK, 0, s>>>(v, pts, n, tol, out, weights, perm, in_order, pairs, window_cursor);
        else k_locate_points<MAXV, false, 4, true><<<grid, BLOCK, 0, s>>>(v, pts, n, tol, out, weights, perm, in_order, pairs, window_cursor);
    } else if (weights)
        k_locate_points<MAXV, true, 8, false><<<grid, BLOCK, 0, s>>>(v, pts, n, tol, out, weights, perm, in_order, pairs, window_cursor);
    else if (MAXV <= 4)
        k_locate_points<(MAXV <= 4 ? MAXV : 4), false, 9, false><<<grid, BLOCK, 0, s>>>(v, pts, n, tol, out, weights, perm, in_order, pairs, window_cursor);
    else
        k_locate_points<MAXV, false, 8, false><<<grid, BLOCK, 0, s>>>(v, pts, n, tol, out, weights, perm, in_order, pairs, window_cursor);
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
    DeepScope deep;
    CT_CHECK(deep.init(tree, n, s));
    CT_CHECK(deep.next_launch());
    v.deep = deep.view;
    PhaseEvents *ev = profile ? phase_events() : nullptr;
    if (ev) CT_CUDA(cudaEventRecord(ev->start, s));
    const bool binned = sort_bits_for(tree, n) > 0 && v.deep.slab == nullptr;  // deep trees: the simple kernel
    if (!binned) {
        // the caller's order: small batches, small trees
        if (ev) CT_CUDA(cudaEventRecord(ev->ordered, s));
        int status = CT_OK;
        if (tree->kind == CT_KIND_EDGES) {
            if (v.deep.slab) k_locate_points_on_edge<true><<<grid_for(n, BLOCK), BLOCK, 0, s>>>(v, pts, n, tol, out, nullptr, nullptr);
            else k_locate_points_on_edge<false><<<grid_for(n, BLOCK), BLOCK, 0, s>>>(v, pts, n, tol, out, nullptr, nullptr);
            CT_LAUNCH_CHECK();
        } else if (tree->M == 3) status = launch_locate_points<3>(v, pts, n, tol, out, weights, nullptr, nullptr, s);
        else if (tree->M == 4) status = launch_locate_points<4>(v, pts, n, tol, out, weights, nullptr, nullptr, s);
        else if (tree->M <= 8) status = launch_locate_points<8>(v, pts, n, tol, out, weights, nullptr, nullptr, s);
        else status = launch_locate_points<32>(v, pts, n, tol, out, weights, nullptr, nullptr, s);
        if (ev) CT_CUDA(cudaEventRecord(ev->done, s));
        return status == CT_OK ? deep.finish() : status;
    }
    // Execution order (CELLTREE_ORDER, measured on C2 with 100 M points): "slabs" (default) appends every point to the slab
    // of its Z-order bin, no counting pass (binning.cuh); "bins" counts first and packs the bins densely (+0.7 ms); "sort"
    // radix-sorts (bin, index) pairs and lets the tiles gather (the sort is 1.3 ms cheaper than "bins", the gathering
    // traversal 1.7 ms dearer); "morton" is round 1's full 24-bit sort of (key, index) pairs with one gathered query per
    // thread (+0.6 ms against "bins" with the window queues).
    static int order_mode = -1;
    if (order_mode < 0) {
        const char *e = getenv("CELLTREE_ORDER");
        order_mode = !e ? 3 : (e[0] == 'b' ? 0 : (e[0] == 's' && e[1] == 'o' ? 1 : (e[0] == 'm' ? 2 : 3)));
    }
    if (order_mode == 2 && tree->kind == CT_KIND_FACES) {
        // full Z-order sort of (key, index) pairs, one query per thread with the point gathered through the permutation
        MortonOrder order;
        CT_CHECK(order.build<KEY_POINT>(tree, reinterpret_cast<const double *>(pts), n, s));
        const int64_t n_windows = (n + WINDOW - 1) >> WINDOW_BITS;
        Scratch<uint2> pairs;
        Scratch<uint32_t> window_cursor;
        const bool queued = n > direct_out_limit();
        if (queued) {
            CT_CHECK(pairs.alloc(n, s));
            CT_CHECK(window_cursor.alloc(n_windows, s));
            CT_CUDA(cudaMemsetAsync(window_cursor.p, 0, n_windows * sizeof(uint32_t), s));
        }
        if (ev) CT_CUDA(cudaEventRecord(ev->ordered, s));
        int status;
        if (tree->M == 3) status = launch_locate_points<3>(v, pts, n, tol, out, weights, order.perm, nullptr, s, pairs.p, window_cursor.p);
        else if (tree->M == 4) status = launch_locate_points<4>(v, pts, n, tol, out, weights, order.perm, nullptr, s, pairs.p, window_cursor.p);
        else if (tree->M <= 8) status = launch_locate_points<8>(v, pts, n, tol, out, weights, order.perm, nullptr, s, pairs.p, window_cursor.p);
        else status = launch_locate_points<32>(v, pts, n, tol, out, weights, order.perm, nullptr, s, pairs.p, window_cursor.p);
        if (ev) CT_CUDA(cudaEventRecord(ev->done, s));
        CT_CHECK(status);
        if (queued) {
            CT_CUDA(cudaFuncSetAttribute(k_windows_to_out, cudaFuncAttributeMaxDynamicSharedMemorySize, WINDOW * (int)sizeof(int32_t)));
            k_windows_to_out<<<(unsigned)n_windows, WINDOW_THREADS, WINDOW * sizeof(int32_t), s>>>(pairs.p, n, out);
            CT_LAUNCH_CHECK();
        }
        return deep.finish();
    }
    PointBins bins;
    BinSort sorted;
    PointSlabs slabs;
    TileInput in{nullptr, nullptr, pts, BinGrid{tree->bbox[0], tree->bbox[2], tree->grid_sx, tree->grid_sy}, nullptr, SlabPlan{}, tree->cells_per_leaf > 2};
    if (order_mode == 3 && n > ((int64_t)400 << 20)) order_mode = 0;  // the slabs of that many points: too much memory
    if (order_mode == 3) {
        CT_CHECK(slabs.build(tree, pts, n, s));
        in.records = slabs.records.p;
        in.slab_fill = slabs.cursor.p;
        in.plan = slabs.plan;
    } else if (order_mode == 0) {
        CT_CHECK(bins.build(tree, pts, n, s));
        in.records = bins.records.p;
    } else {
        CT_CHECK(sorted.build(in.grid, pts, n, s));
        in.perm = sorted.perm;
    }
    Scratch<uint32_t> window_cursor;
    const int64_t n_windows = (n + WINDOW - 1) >> WINDOW_BITS;
    Scratch<uint2> pairs;
    const bool queued = n > direct_out_limit();
    if (queued) {
        CT_CHECK(pairs.alloc(n, s));
        CT_CHECK(window_cursor.alloc(n_windows, s));
        CT_CUDA(cudaMemsetAsync(window_cursor.p, 0, n_windows * sizeof(uint32_t), s));
    }
    if (ev) CT_CUDA(cudaEventRecord(ev->ordered, s));
    int status;
    uint32_t *wc = window_cursor.p;
    if (tree->kind == CT_KIND_EDGES) status = launch_locate_points_binned<0>(v, in, n, tol, pairs.p, wc, out, nullptr, s);
    else if (tree->M == 3) status = launch_locate_points_binned<3>(v, in, n, tol, pairs.p, wc, out, weights, s);
    else if (tree->M == 4) status = launch_locate_points_binned<4>(v, in, n, tol, pairs.p, wc, out, weights, s);
    else if (tree->M <= 8) status = launch_locate_points_binned<8>(v, in, n, tol, pairs.p, wc, out, weights, s);
    else status = launch_locate_points_binned<32>(v, in, n, tol, pairs.p, wc, out, weights, s);
    if (status == CT_OK && in.slab_fill) {
        const uint32_t *list = slabs.overflow.p, *count = slabs.overflow_count();
        if (tree->kind == CT_KIND_EDGES) status = launch_locate_points_overflow<0>(v, pts, list, count, tol, pairs.p, wc, out, nullptr, s);
        else if (tree->M == 3) status = launch_locate_points_overflow<3>(v, pts, list, count, tol, pairs.p, wc, out, weights, s);
        else if (tree->M == 4) status = launch_locate_points_overflow<4>(v, pts, list, count, tol, pairs.p, wc, out, weights, s);
        else if (tree->M <= 8) status = launch_locate_points_overflow<8>(v, pts, list, count, tol, pairs.p, wc, out, weights, s);
        else status = launch_locate_points_overflow<32>(v, pts, list, count, tol, pairs.p, wc, out, weights, s);
    }
    if (ev) CT_CUDA(cudaEventRecord(ev->done, s));
    if (status == CT_OK && queued) {
        CT_CUDA(cudaFuncSetAttribute(k_windows_to_out, cudaFuncAttributeMaxDynamicSharedMemorySize, WINDOW * (int)sizeof(int32_t)));
        k_windows_to_out<<<(unsigned)n_windows, WINDOW_THREADS, WINDOW * sizeof(int32_t), s>>>(pairs.p, n, out);
        CT_LAUNCH_CHECK();
    }
    return status == CT_OK ? deep.finish() : status;
}

}  // namespace ct

using namespace ct;

// after a call that computed weights: did a kernel divide by zero where the reference raises?
static int check_zero_division(cudaStream_t s) {
    int flag = 0;
    CT_CUDA(cudaMemcpyFromSymbolAsync(&flag, g_zero_division, sizeof(int), 0, cudaMemcpyDeviceToHost, s));
    CT_CUDA(cudaStreamSynchronize(s));
    if (!flag) return CT_OK;
    const int zero = 0;
    CT_CUDA(cudaMemcpyToSymbolAsync(g_zero_division, &zero, sizeof(int), 0, cudaMemcpyHostToDevice, s));
    CT_CUDA(cudaStreamSynchronize(s));
    set_error("division by zero");
    return CT_ERR_ZERO_DIVISION;
}

static int locate_points_entry(const ct_tree *tree, const double *points, int64_t n, double tolerance, int64_t *out_index, double *weights,
                               int32_t mem);

extern "C" int ct_locate_points(const ct_tree *tree, const double *points, int64_t n, double tolerance, int64_t *out_index,
                                double *weights, int32_t mem) {
    const int status = locate_points_entry(tree, points, n, tolerance, out_index, weights, mem);
    if (status != CT_OK || !weights || n == 0) return status;
    CT_ON_DEVICE(tree->device);
    return check_zero_division(current_stream());
}

static int locate_points_entry(const ct_tree *tree, const double *points, int64_t n, double tolerance, int64_t *out_index, double *weights,
                               int32_t mem) {
    if (!tree || n < 0 || (n > 0 && (!points || !out_index))) {
        set_error("ct_locate_points: null argument");
        return CT_ERR_VALUE;
    }
    if (weights && tree->kind != CT_KIND_FACES) {
        set_error("ct_locate_points: barycentric weights need a face tree");
        return CT_ERR_VALUE;
    }
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

// ---- diagnostics for the roofline line of bench.py -------------------------------------------------------------------
namespace ct {
template <int MAXV>
__global__ void __launch_bounds__(BLOCK) k_locate_points_stats(TreeView t, const double2 *__restrict__ points, int64_t n, double tolerance,
                                                               unsigned long long *__restrict__ stats) {
    const int64_t i = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
    CountingProbe probe;
    unsigned found = 0;
    if (i < n) {
        const double2 pt = points[i];
        found = locate_point<MAXV, CountingProbe, true>(t, P2{pt.x, pt.y}, tolerance, &probe) >= 0 ? 1u : 0u;
    }
    const unsigned v[6] = {probe.slots, probe.headers, probe.cells, probe.pushes, probe.entries, found};
#pragma unroll
    for (int k = 0; k < 6; k++) {
        const unsigned sum = __reduce_add_sync(0xffffffffu, v[k]);
        if ((threadIdx.x & 31) == 0 && sum) atomicAdd(stats + k, (unsigned long long)sum);
    }
}

template <int BYTES>
__global__ void __launch_bounds__(256) k_read_sweep(const uint4 *__restrict__ data, size_t n_words, int repeats, unsigned *__restrict__ sink) {
    // every thread reads 16-byte words at a grid stride, `repeats` sweeps over the buffer
    unsigned acc = 0;
    const size_t stride = (size_t)gridDim.x * 256;
    for (int r = 0; r < repeats; r++)
        for (size_t w = (size_t)blockIdx.x * 256 + threadIdx.x; w < n_words; w += stride) {
            uint4 q;
            asm volatile("ld.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w) : "l"(data + w));
            acc += q.x ^ q.y ^ q.z ^ q.w;
        }
    if (acc == 0x9e3779b9u) *sink = acc;  // keeps the loads alive
}
}  // namespace ct

// What one launch of the point traversal touches, summed over the n queries (points on the device, caller's order):
// stats[0] node slots read (16 B each), [1] treelet headers read (16 B), [2] cells tested (each: one row of elem_xy,
// 16 B per vertex, and one row of elements), [3] siblings deferred to the stack, [4] queries that started from the
// entry grid rather than the root, [5] queries that found a cell.
extern "C" int ct_locate_points_stats(const ct_tree *tree, const double *points, int64_t n, double tolerance, int64_t *stats) {
    if (!tree || !points || n <= 0 || !stats || tree->kind != CT_KIND_FACES) {
        set_error("ct_locate_points_stats: bad argument (needs a face tree and device points)");
        return CT_ERR_VALUE;
    }
    CT_ON_DEVICE(tree->device);
    cudaStream_t s = current_stream();
    TreeView v = tree->view();
    DeepScope deep;
    CT_CHECK(deep.init(tree, n, s));
    CT_CHECK(deep.next_launch());
    v.deep = deep.view;
    Scratch<unsigned long long> d_stats;
    CT_CHECK(d_stats.alloc(6, s));
    CT_CUDA(cudaMemsetAsync(d_stats.p, 0, 6 * sizeof(unsigned long long), s));
    const double2 *pts = reinterpret_cast<const double2 *>(points);
    const int grid = grid_for(n, BLOCK);
    if (tree->M == 3) k_locate_points_stats<3><<<grid, BLOCK, 0, s>>>(v, pts, n, tolerance, d_stats.p);
    else if (tree->M == 4) k_locate_points_stats<4><<<grid, BLOCK, 0, s>>>(v, pts, n, tolerance, d_stats.p);
    else if (tree->M <= 8) k_locate_points_stats<8><<<grid, BLOCK, 0, s>>>(v, pts, n, tolerance, d_stats.p);
    else k_locate_points_stats<32><<<grid, BLOCK, 0, s>>>(v, pts, n, tolerance, d_stats.p);
    CT_LAUNCH_CHECK();
    unsigned long long h[6];
    CT_CUDA(cudaMemcpyAsync(h, d_stats.p, sizeof(h), cudaMemcpyDeviceToHost, s));
    CT_CUDA(cudaStreamSynchronize(s));
    for (int k = 0; k < 6; k++) stats[k] = (int64_t)h[k];
    return deep.finish();
}

// Read bandwidth of this GPU measured in place: `repeats` sweeps of 16-byte loads over a buffer of `bytes` bytes by a
// grid that fills the machine.  A buffer well inside the L2 (64 MB of 126 MB) gives the L2 read rate after the first
// sweep has brought it in; a buffer many times the L2 gives the HBM read rate.  The first sweep is not timed.
extern "C" int ct_measure_read_bandwidth(size_t bytes, int32_t repeats, double *gb_per_s) {
    if (bytes < 4096 || repeats < 1 || !gb_per_s) {
        set_error("ct_measure_read_bandwidth: bad argument");
        return CT_ERR_VALUE;
    }
    cudaStream_t s = current_stream();
    Scratch<uint4> buffer;
    Scratch<unsigned> sink;
    const size_t words = bytes / 16;
    CT_CHECK(buffer.alloc(words, s));
    CT_CHECK(sink.alloc(1, s));
    CT_CUDA(cudaMemsetAsync(buffer.p, 1, words * 16, s));
    int device = 0, sms = 0;
    CT_CUDA(cudaGetDevice(&device));
    CT_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    const int grid = sms * 8;
    cudaEvent_t e0, e1;
    CT_CUDA(cudaEventCreate(&e0));
    CT_CUDA(cudaEventCreate(&e1));
    k_read_sweep<16><<<grid, 256, 0, s>>>(buffer.p, words, 1, sink.p);
    CT_CUDA(cudaEventRecord(e0, s));
    k_read_sweep<16><<<grid, 256, 0, s>>>(buffer.p, words, repeats, sink.p);
    CT_CUDA(cudaEventRecord(e1, s));
    CT_CUDA(cudaEventSynchronize(e1));
    count_launch(2);
    float ms = 0.f;
    CT_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *gb_per_s = (double)words * 16.0 * repeats / (ms * 1e-3) / 1e9;
    return CT_OK;
}
