// binning.cuh -- spatial binning of point queries: the execution order of ct_locate_points on large batches.
// (The default are the SLAB bins at the end of this file; the counted bins described first are CELLTREE_ORDER=bins.)
//
// Queries are independent (query.py:110-117 is a prange), so which thread handles which query is an execution detail.
// What the traversal needs from the order is (a) that the leaves a block touches are few, so that tree data comes out
// of DRAM once, and (b) that the 32 lanes of a warp sit in neighbouring leaves.  A full radix sort of (key, index)
// pairs delivers both but costs three passes plus a 16-byte gather per query that drags a whole DRAM atom along,
// and a fourth pass to put the results back (round 1: 3.9 of the 7.3 ms of a 100 M-point step on C2).
//
// Here the POINTS are moved, once:
//   k_bin_count    histogram of the 16-bit coarse Z-order key (256 x 256 cells over the tree's bounding box)
//   k_bin_offsets  exclusive scan of the 65 536 counts (one block)
//   k_bin_scatter  every point is appended to its bin as one 32-byte record {x, y, index, 24-bit key}: the position
//                  comes from an atomic on the bin's cursor (lanes of a warp with equal keys share one), the record
//                  is one 256-bit store; a bin's open 128-byte line stays in L2 until it is full (65 536 open lines =
//                  8 MB), so DRAM sees whole lines.  The order inside a bin is whatever the atomics gave: it does not
//                  matter, because ...
//   the traversal  (points.cu) takes TILES of 2048 consecutive records, sorts each tile in shared memory by the
//                  next 8 key bits (cub::BlockRadixSort over as many bits as the tile spans) and walks the tree in
//                  that order: the lanes of a warp are as close as after a full 24-bit sort.
// Results go back without a sort as well: a thread appends (index, result) to the queue of the index's WINDOW
// (16 384 consecutive queries; again an atomic cursor + an 8-byte store that L2 merges), and k_windows_to_out turns
// every queue into its slice of `out` through shared memory, written as whole lines.  Every index occurs exactly once,
// so window w's queue is exactly slots [w << 14, (w + 1) << 14) of one n-element array: no histogram is needed.
#pragma once

#include "morton.cuh"

namespace ct {

struct __align__(32) PointRecord {
    double x, y;
    uint32_t index;  // position of the query in the caller's array
    uint32_t key;    // 24-bit Z-order key over the tree's bounding box (top 16 bits = the bin)
    uint64_t pad;
};
static_assert(sizeof(PointRecord) == 32, "one record = one 32-byte sector");

constexpr int BIN_BITS = 16;
constexpr int N_BINS = 1 << BIN_BITS;
constexpr int FINE_BITS = 8;  // key bits below the bin that the tiles are sorted by
constexpr int WINDOW_BITS = 14;
constexpr int WINDOW = 1 << WINDOW_BITS;

struct BinGrid {
    double xmin, ymin, sx, sy;
};

__device__ __forceinline__ uint32_t point_key24(const BinGrid &g, double x, double y) {
    const uint32_t ix = grid_coord(x, g.xmin, g.sx), iy = grid_coord(y, g.ymin, g.sy);
    return (spread16(ix) | (spread16(iy) << 1)) >> 8;
}

// 256-bit global accesses (sm_100: LDG.E.ENL2.256 / STG.E.ENL2.256): one instruction per record
__device__ __forceinline__ void store_record(PointRecord *dst, double x, double y, uint32_t index, uint32_t key) {
    asm volatile("st.global.v4.b64 [%0], {%1, %2, %3, %4};" ::"l"(dst), "l"(__double_as_longlong(x)), "l"(__double_as_longlong(y)),
                 "l"((unsigned long long)index | ((unsigned long long)key << 32)), "l"(0ULL)
                 : "memory");
}
__device__ __forceinline__ void load_record(const PointRecord *src, double &x, double &y, uint32_t &index, uint32_t &key) {
    unsigned long long a, b, c, d;
    asm volatile("ld.global.nc.L1::no_allocate.v4.b64 {%0, %1, %2, %3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(src));
    (void)d;
    x = __longlong_as_double((long long)a);
    y = __longlong_as_double((long long)b);
    index = (uint32_t)c;
    key = (uint32_t)(c >> 32);
}

constexpr int BIN_BLOCK = 256;
// Measurement switches (build_ext --variant TAG -DNAME=VALUE; DESIGN.md 4.2 quotes what they showed), all off by default:
//   CT_CURSOR_STRIDE  words between the cursors of the counted bins (8 / 32: one cursor per sector / line -- no effect)
//   CT_EXP            1: k_bin_scatter without its stores, 3: without its atomics (hashed positions; results unusable)
#ifndef CT_CURSOR_STRIDE
#define CT_CURSOR_STRIDE 1
#endif
#ifndef CT_EXP
#define CT_EXP 0
#endif
#ifndef CT_BIN_PER_THREAD
#define CT_BIN_PER_THREAD 8
#endif
constexpr int BIN_PER_THREAD = CT_BIN_PER_THREAD;

// One atomic per RUN of equal keys among consecutive lanes of a warp: sorted, gridded or clustered inputs would otherwise
// serialise on a few counters (random inputs have no runs and pay one vote + one shuffle for the check).  Returns
// this lane's rank within its run, the run's length and its first lane.
__device__ __forceinline__ void warp_runs(uint32_t key, uint32_t &rank, uint32_t &size, int &leader) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t before = __shfl_up_sync(0xffffffffu, key, 1);
    const uint32_t heads = __ballot_sync(0xffffffffu, lane == 0 || key != before);  // bit l: lane l starts a run
    leader = 31 - __clz(heads & (0xffffffffu >> (31u - lane)));                     // the last head at or below this lane
    const uint32_t later = lane == 31 ? 0u : (heads & (0xfffffffeu << lane));       // heads above this lane
    const int end = later ? __ffs(later) - 1 : 32;
    rank = lane - (uint32_t)leader;
    size = (uint32_t)(end - leader);
}

static __global__ void __launch_bounds__(BIN_BLOCK) k_bin_count(const double2 *__restrict__ points, int64_t n, BinGrid g,
                                                                 uint32_t *__restrict__ count) {
    const int64_t first = (int64_t)blockIdx.x * (BIN_BLOCK * BIN_PER_THREAD) + threadIdx.x;
    double2 p[BIN_PER_THREAD];
#pragma unroll
    for (int k = 0; k < BIN_PER_THREAD; k++) {
        const int64_t i = first + (int64_t)k * BIN_BLOCK;
        p[k] = i < n ? __ldcs(points + i) : make_double2(0.0, 0.0);
    }
#pragma unroll
    for (int k = 0; k < BIN_PER_THREAD; k++) {
        const int64_t i = first + (int64_t)k * BIN_BLOCK;
        // lanes past the end form runs of their own key (N_BINS: no bin) and count nothing
        const uint32_t bin = i < n ? point_key24(g, p[k].x, p[k].y) >> FINE_BITS : (uint32_t)N_BINS;
        uint32_t rank, size;
        int leader;
        warp_runs(bin, rank, size, leader);
        if (rank == 0 && bin < (uint32_t)N_BINS) atomicAdd(count + bin * CT_CURSOR_STRIDE, size);
    }
}

// exclusive scan of the N_BINS counts, in place (one block of 1024 threads, 64 bins each)
static __global__ void __launch_bounds__(1024) k_bin_offsets(uint32_t *__restrict__ count) {
    __shared__ uint32_t warp_sum[32];
    constexpr int PER = N_BINS / 1024;
    uint4 *src = reinterpret_cast<uint4 *>(count + threadIdx.x * PER);
    uint32_t total = 0;
#pragma unroll 4
    for (int k = 0; k < PER / 4; k++) {
        const uint4 q = src[k];
        total += q.x + q.y + q.z + q.w;
    }
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t incl = total;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t up = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= (uint32_t)d) incl += up;
    }
    if (lane == 31) warp_sum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = warp_sum[lane], wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t up = __shfl_up_sync(0xffffffffu, wi, d);
            if (lane >= (uint32_t)d) wi += up;
        }
        warp_sum[lane] = wi - w;
    }
    __syncthreads();
    uint32_t running = warp_sum[warp] + incl - total;
#pragma unroll 4
    for (int k = 0; k < PER / 4; k++) {
        const uint4 c = src[k];  // this thread's own 64 counts again (L1)
        uint4 q;
        q.x = running, running += c.x;
        q.y = running, running += c.y;
        q.z = running, running += c.z;
        q.w = running, running += c.w;
        src[k] = q;
    }
}

#if CT_CURSOR_STRIDE != 1
static __global__ void __launch_bounds__(1024) k_bin_offsets_strided(uint32_t *__restrict__ count) {  // experiment: slow and simple
    __shared__ uint32_t part[1024];
    constexpr int PER = N_BINS / 1024;
    uint32_t total = 0;
    for (int k = 0; k < PER; k++) total += count[(threadIdx.x * PER + k) * CT_CURSOR_STRIDE];
    part[threadIdx.x] = total;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t run = 0;
        for (int t = 0; t < 1024; t++) { uint32_t v = part[t]; part[t] = run; run += v; }
    }
    __syncthreads();
    uint32_t running = part[threadIdx.x];
    for (int k = 0; k < PER; k++) {
        uint32_t *c = count + (threadIdx.x * PER + k) * CT_CURSOR_STRIDE;
        const uint32_t v = *c;
        *c = running;
        running += v;
    }
}
#endif

static __global__ void __launch_bounds__(BIN_BLOCK) k_bin_scatter(const double2 *__restrict__ points, int64_t n, BinGrid g,
                                                                   uint32_t *__restrict__ cursor, PointRecord *__restrict__ records) {
    const int64_t first = (int64_t)blockIdx.x * (BIN_BLOCK * BIN_PER_THREAD) + threadIdx.x;
    double2 p[BIN_PER_THREAD];
#pragma unroll
    for (int k = 0; k < BIN_PER_THREAD; k++) {
        const int64_t i = first + (int64_t)k * BIN_BLOCK;
        p[k] = i < n ? __ldcs(points + i) : make_double2(0.0, 0.0);
    }
    // all the atomics of a thread are issued before the first result is needed: BIN_PER_THREAD round trips in flight
    uint32_t base[BIN_PER_THREAD], key[BIN_PER_THREAD], rank[BIN_PER_THREAD];
    int leader[BIN_PER_THREAD];
#pragma unroll
    for (int k = 0; k < BIN_PER_THREAD; k++) {
        const int64_t i = first + (int64_t)k * BIN_BLOCK;
        key[k] = point_key24(g, p[k].x, p[k].y);
        const uint32_t bin = i < n ? key[k] >> FINE_BITS : (uint32_t)N_BINS;
        uint32_t size;
        warp_runs(bin, rank[k], size, leader[k]);
        base[k] = 0;
#if CT_EXP == 3
        base[k] = (uint32_t)(((uint64_t)i * 2654435761ull) % (uint64_t)n);
        rank[k] = 0; leader[k] = threadIdx.x & 31;
#else
        if (rank[k] == 0 && bin < (uint32_t)N_BINS) base[k] = atomicAdd(cursor + bin * CT_CURSOR_STRIDE, size);
#endif
    }
#pragma unroll
    for (int k = 0; k < BIN_PER_THREAD; k++) {
        const int64_t i = first + (int64_t)k * BIN_BLOCK;
        const uint32_t pos = __shfl_sync(0xffffffffu, base[k], leader[k]) + rank[k];
#if CT_EXP == 1
        if (i < n && pos == 0xffffffffu) store_record(records, p[k].x, p[k].y, (uint32_t)i, key[k]);
#else
        if (i < n) store_record(records + pos, p[k].x, p[k].y, (uint32_t)i, key[k]);
#endif
    }
}

// queue of window w (slots [w << WINDOW_BITS, ...) of `pairs`, any order) -> out[w << WINDOW_BITS ...], as int64
constexpr int WINDOW_THREADS = 512;
static __global__ void __launch_bounds__(WINDOW_THREADS) k_windows_to_out(const uint2 *__restrict__ pairs, int64_t n, int64_t *__restrict__ out) {
    extern __shared__ int32_t s_result[];
    const int64_t lo = (int64_t)blockIdx.x << WINDOW_BITS;
    const int m = (int)((n - lo) < WINDOW ? (n - lo) : WINDOW);
    for (int j = threadIdx.x; j < m; j += WINDOW_THREADS) {
        const uint2 pr = __ldcs(pairs + lo + j);
        s_result[pr.x & (WINDOW - 1)] = (int32_t)pr.y;
    }
    __syncthreads();
    if ((m & 1) == 0 && ((reinterpret_cast<uintptr_t>(out + lo) & 15) == 0)) {
        longlong2 *o = reinterpret_cast<longlong2 *>(out + lo);
        for (int j = threadIdx.x; j < m / 2; j += WINDOW_THREADS)
            __stcs(o + j, make_longlong2((long long)s_result[2 * j], (long long)s_result[2 * j + 1]));
    } else {
        for (int j = threadIdx.x; j < m; j += WINDOW_THREADS) __stcs(out + lo + j, (int64_t)s_result[j]);
    }
}

// The binned copy of a batch of points.
struct PointBins {
    Scratch<PointRecord> records;
    Scratch<uint32_t> cursor;  // N_BINS bin cursors

    int build(const ct_tree *tree, const double2 *points, int64_t n, cudaStream_t s) {
        CT_CHECK(records.alloc(n, s));
        CT_CHECK(cursor.alloc(N_BINS * CT_CURSOR_STRIDE, s));
        CT_CUDA(cudaMemsetAsync(cursor.p, 0, (N_BINS * CT_CURSOR_STRIDE) * sizeof(uint32_t), s));
        const BinGrid g{tree->bbox[0], tree->bbox[2], tree->grid_sx, tree->grid_sy};
        const int grid = grid_for(n, BIN_BLOCK * BIN_PER_THREAD);
        k_bin_count<<<grid, BIN_BLOCK, 0, s>>>(points, n, g, cursor.p);
        CT_LAUNCH_CHECK();
#if CT_CURSOR_STRIDE == 1
        k_bin_offsets<<<1, 1024, 0, s>>>(cursor.p);
#else
        k_bin_offsets_strided<<<1, 1024, 0, s>>>(cursor.p);
#endif
        CT_LAUNCH_CHECK();
        k_bin_scatter<<<grid, BIN_BLOCK, 0, s>>>(points, n, g, cursor.p, records.p);
        CT_LAUNCH_CHECK();
        return CT_OK;
    }
};

// ---- slab bins: the same bins without the counting pass ------------------------------------------------------------------
// Every bin owns a slab of SLAB records; a point takes the next slot of its bin's slab (the same atomic as above), and
// there are so many bins that a slab is seven eighths full on average (uniform points: 1792 +- 42 per bin, the slab's
// 2048 is six standard deviations away; the few points beyond it take the overflow path).  No histogram, no offsets: the counting
// pass (0.7 ms of a 100 M-point step) is gone, and a tile of the traversal is exactly one bin, sorted by the key bits
// below the bin alone.  A point that finds its slab full -- crowded query sets -- is noted by index in an overflow list
// and walked by k_locate_points_overflow in arrival order (such points are neighbours in space: their part of the tree
// is small).  Nothing depends on the host knowing how many there were.
constexpr int SLAB = 2048;

struct SlabPlan {
    uint32_t bins;  // any number, not only powers of two: bin = key24 * bins >> 24 maps the Z-order curve onto them in order
    int shift;      // a tile is sorted by (key24 - first key of its bin) >> shift ...
    int sort_bits;  // ... which has this many bits (<= MAX_SORT_BITS: the counters of the tile's counting sort, tile_order.cuh)
    static constexpr int MAX_SORT_BITS = 10;
    static SlabPlan make(int64_t n) {
        SlabPlan p;
        const int64_t target = SLAB * 7 / 8;  // mean points per bin: 1792 +- 42 for uniform points, 6 deviations below the slab
        int64_t bins = (n + target - 1) / target;
        if (bins < 1) bins = 1;
        if (bins > (1 << 22)) bins = 1 << 22;
        p.bins = (uint32_t)bins;
        const uint32_t range = (uint32_t)(((uint64_t)1 << 24) / p.bins) + 2;  // keys per bin, at most
        int bits = 1;
        while (bits < 24 && (1u << bits) < range) bits++;
        p.sort_bits = bits < MAX_SORT_BITS ? bits : MAX_SORT_BITS;
        p.shift = bits - p.sort_bits;
        return p;
    }
};
__device__ __forceinline__ uint32_t slab_bin(uint32_t key24, uint32_t bins) { return (uint32_t)(((uint64_t)key24 * bins) >> 24); }
__device__ __forceinline__ uint32_t slab_first_key(uint32_t bin, uint32_t bins) {  // the smallest key24 that slab_bin() sends to `bin`
    return (uint32_t)((((uint64_t)bin << 24) + bins - 1) / bins);
}

static __global__ void __launch_bounds__(BIN_BLOCK) k_slab_scatter(const double2 *__restrict__ points, int64_t n, BinGrid g, uint32_t bins,
                                                                    uint32_t *__restrict__ cursor, PointRecord *__restrict__ records,
                                                                    uint32_t *__restrict__ overflow, uint32_t *__restrict__ overflow_count) {
    const int64_t first = (int64_t)blockIdx.x * (BIN_BLOCK * BIN_PER_THREAD) + threadIdx.x;
    double2 p[BIN_PER_THREAD];
#pragma unroll
    for (int k = 0; k < BIN_PER_THREAD; k++) {
        const int64_t i = first + (int64_t)k * BIN_BLOCK;
        p[k] = i < n ? __ldcs(points + i) : make_double2(0.0, 0.0);
    }
    uint32_t base[BIN_PER_THREAD], key[BIN_PER_THREAD], rank[BIN_PER_THREAD], bin[BIN_PER_THREAD];
    int leader[BIN_PER_THREAD];
#pragma unroll
    for (int k = 0; k < BIN_PER_THREAD; k++) {
        const int64_t i = first + (int64_t)k * BIN_BLOCK;
        key[k] = point_key24(g, p[k].x, p[k].y);
        bin[k] = i < n ? slab_bin(key[k], bins) : 0xffffffffu;
        uint32_t size;
        warp_runs(bin[k], rank[k], size, leader[k]);
        base[k] = 0;
        if (rank[k] == 0 && bin[k] != 0xffffffffu) base[k] = atomicAdd(cursor + bin[k], size);
    }
#pragma unroll
    for (int k = 0; k < BIN_PER_THREAD; k++) {
        const int64_t i = first + (int64_t)k * BIN_BLOCK;
        const uint32_t slot = __shfl_sync(0xffffffffu, base[k], leader[k]) + rank[k];
        if (i >= n) continue;
        if (slot < (uint32_t)SLAB) {
            store_record(records + ((size_t)bin[k] * SLAB + slot), p[k].x, p[k].y, (uint32_t)i, key[k]);
        } else {
            overflow[atomicAdd(overflow_count, 1u)] = (uint32_t)i;
        }
    }
}

struct PointSlabs {
    Scratch<PointRecord> records;  // bins * SLAB
    Scratch<uint32_t> cursor;      // bins cursors (after the scatter: how many points each bin received), then the overflow count
    Scratch<uint32_t> overflow;    // indices of the points that found their slab full
    SlabPlan plan;

    int build(const ct_tree *tree, const double2 *points, int64_t n, cudaStream_t s) {
        plan = SlabPlan::make(n);
        const int64_t bins = plan.bins;
        CT_CHECK(records.alloc((size_t)bins * SLAB, s));
        CT_CHECK(cursor.alloc(bins + 1, s));
        CT_CHECK(overflow.alloc(n, s));
        CT_CUDA(cudaMemsetAsync(cursor.p, 0, (bins + 1) * sizeof(uint32_t), s));
        const BinGrid g{tree->bbox[0], tree->bbox[2], tree->grid_sx, tree->grid_sy};
        k_slab_scatter<<<grid_for(n, BIN_BLOCK * BIN_PER_THREAD), BIN_BLOCK, 0, s>>>(points, n, g, plan.bins, cursor.p, records.p,
                                                                                    overflow.p, cursor.p + bins);
        CT_LAUNCH_CHECK();
        return CT_OK;
    }
    const uint32_t *overflow_count() const { return cursor.p + plan.bins; }
};

// The other way to the same execution order: (bin, index) pairs sorted by the 16-bit bin (two 8-bit passes of CUB's radix
// sort on 2-byte keys), the points gathered by the traversal's tiles.
static __global__ void __launch_bounds__(256) k_bin_keys(const double2 *__restrict__ points, int64_t n, BinGrid g, uint16_t *__restrict__ keys,
                                                          uint32_t *__restrict__ index) {
    const int64_t first = (int64_t)blockIdx.x * 1024 + threadIdx.x;
    double2 p[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int64_t i = first + k * 256;
        p[k] = i < n ? __ldcs(points + i) : make_double2(0.0, 0.0);
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int64_t i = first + k * 256;
        if (i < n) {
            keys[i] = (uint16_t)(point_key24(g, p[k].x, p[k].y) >> FINE_BITS);
            index[i] = (uint32_t)i;
        }
    }
}

struct BinSort {
    Scratch<uint16_t> keys_a, keys_b;
    Scratch<uint32_t> idx_a, idx_b;
    Scratch<char> tmp;
    const uint32_t *perm = nullptr;

    int build(const BinGrid &g, const double2 *points, int64_t n, cudaStream_t s) {
        CT_CHECK(keys_a.alloc(n, s));
        CT_CHECK(keys_b.alloc(n, s));
        CT_CHECK(idx_a.alloc(n, s));
        CT_CHECK(idx_b.alloc(n, s));
        k_bin_keys<<<grid_for(n, 1024), 256, 0, s>>>(points, n, g, keys_a.p, idx_a.p);
        CT_LAUNCH_CHECK();
        cub::DoubleBuffer<uint16_t> d_keys(keys_a.p, keys_b.p);
        cub::DoubleBuffer<uint32_t> d_vals(idx_a.p, idx_b.p);
        size_t bytes = 0;
        CT_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, d_keys, d_vals, n, 0, BIN_BITS, s));
        CT_CHECK(tmp.alloc(bytes, s));
        CT_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, bytes, d_keys, d_vals, n, 0, BIN_BITS, s));
        count_launch(3);
        perm = d_vals.Current();
        return CT_OK;
    }
};

}  // namespace ct
