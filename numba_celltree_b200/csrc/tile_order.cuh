// Order of a tile's queries along the Z-order curve: a counting sort in shared memory.
//
// A tile is one bin of the spatial binning (binning.cuh): up to TILE points whose keys, relative to the bin's first key,
// have only a few bits (<= ORDER_MAX_BITS).  One pass is enough -- a shared-memory atomic per point gives its rank among
// the points of the same key AND counts the key, a block scan of the counters gives every key's first position -- where a
// radix sort pays a ranking, a scan and an exchange of keys and values per 4-bit digit.  Points with equal keys end up in
// whatever order the atomics gave them; the order of execution is invisible in the results.
#pragma once
#include <cstdint>

namespace ct {

constexpr int ORDER_MAX_BITS = 10;

template <int THREADS, int ITEMS>
struct TileOrder {
    static constexpr int TILE = THREADS * ITEMS;
    static constexpr int COUNTERS = 1 << ORDER_MAX_BITS;
    static constexpr int PER_THREAD = COUNTERS / THREADS;
    static_assert(COUNTERS % THREADS == 0 && THREADS % 32 == 0 && THREADS <= 1024, "block shape");
    struct Storage {
        alignas(16) uint32_t count[COUNTERS];  // per key: number of points, then the first sorted position
        uint32_t warp_total[THREADS / 32];
        uint16_t order[TILE];  // sorted position -> place of the point in the tile
    };

    // key[k] (< 1 << bits, bits <= ORDER_MAX_BITS) belongs to the point at place k * THREADS + threadIdx.x; places >= m are
    // empty.  Afterwards s.order[0 .. m) lists the places by key.  Ends with a barrier.
    static __device__ __forceinline__ void sort(Storage &s, const uint32_t (&key)[ITEMS], int m, int bits) {
        const int tid = threadIdx.x;
#pragma unroll
        for (int c = 0; c < PER_THREAD; c++) s.count[c * THREADS + tid] = 0;
        __syncthreads();
        uint32_t rank[ITEMS];
#pragma unroll
        for (int k = 0; k < ITEMS; k++) {
            rank[k] = 0;
            if (k * THREADS + tid < m) rank[k] = atomicAdd(&s.count[key[k]], 1u);
        }
        __syncthreads();
        // exclusive scan of the counters: thread t owns counters [t * PER_THREAD, (t + 1) * PER_THREAD)
        // (unused counters are zero)
        uint32_t mine[PER_THREAD], sum = 0;
        if constexpr (PER_THREAD == 4) {
            const uint4 v = reinterpret_cast<const uint4 *>(s.count)[tid];
            mine[0] = v.x, mine[1] = v.y, mine[2] = v.z, mine[3] = v.w;
        } else {
#pragma unroll
            for (int c = 0; c < PER_THREAD; c++) mine[c] = s.count[tid * PER_THREAD + c];
        }
#pragma unroll
        for (int c = 0; c < PER_THREAD; c++) sum += mine[c];
        uint32_t inclusive = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t up = __shfl_up_sync(0xffffffffu, inclusive, d);
            if ((tid & 31) >= d) inclusive += up;
        }
        if ((tid & 31) == 31) s.warp_total[tid >> 5] = inclusive;
        __syncthreads();
        uint32_t before = inclusive - sum;
#pragma unroll
        for (int w = 0; w < THREADS / 32; w++)
            if (w < (tid >> 5)) before += s.warp_total[w];
        if (tid * PER_THREAD < (1 << bits)) {
            uint32_t first[PER_THREAD];
#pragma unroll
            for (int c = 0; c < PER_THREAD; c++) {
                first[c] = before;
                before += mine[c];
            }
            if constexpr (PER_THREAD == 4) {
                reinterpret_cast<uint4 *>(s.count)[tid] = make_uint4(first[0], first[1], first[2], first[3]);
            } else {
#pragma unroll
                for (int c = 0; c < PER_THREAD; c++) s.count[tid * PER_THREAD + c] = first[c];
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < ITEMS; k++) {
            const int place = k * THREADS + tid;
            if (place < m) s.order[s.count[key[k]] + rank[k]] = (uint16_t)place;
        }
        __syncthreads();
    }
};

}  // namespace ct
