// hitlog.cuh -- single-pass assembly of variable-length results.
//
// "Count, then fill" runs the traversal twice.  Where the per-candidate test is expensive (the segment clips of
// intersect_edges) the first pass instead counts the
// hits of every query AND appends each hit (query, rank within the query, cell[, payload]) to a log, in whatever order
// the warps get there; after the scan of the counts the caller moves every entry into its query's range of the result
// (edges.cu: k_place_sources + k_rank_and_move), which restores the reference's order.  If the log's capacity does not
// suffice, the caller falls back to the second traversal, which writes the pairs in place.
#pragma once

#include "common.cuh"

namespace ct {

struct HitLog {
    unsigned long long *count;  // entries requested so far (may exceed capacity)
    int64_t capacity;
    int32_t *q, *k, *j;
    double *xy;  // 4 doubles per entry, or nullptr
};

// Log position of the calling lane's hit: the lanes that arrive together reserve their entries with one atomic.
__device__ __forceinline__ int64_t hitlog_reserve(const HitLog &log) {
    const unsigned lane = threadIdx.x & 31u;
    const unsigned mask = __activemask();
    const int leader = __ffs(mask) - 1;
    unsigned long long first = 0;
    if ((int)lane == leader) first = atomicAdd(log.count, (unsigned long long)__popc(mask));
    first = __shfl_sync(mask, first, leader);
    return (int64_t)first + __popc(mask & ((1u << lane) - 1u));
}

// Owns the log's device buffers for the duration of a call.
struct HitLogBuffers {
    Scratch<int32_t> q, k, j;
    Scratch<double> xy;
    Scratch<unsigned long long> count;
    HitLog log{};

    // room for `per_query` hits per query on average, at most a quarter of the free device memory
    int alloc(int64_t n, int64_t per_query, bool with_xy, cudaStream_t s) {
        const int64_t entry_bytes = with_xy ? 44 : 12;
        size_t free_bytes = 0, total_bytes = 0;
        CT_CUDA(cudaMemGetInfo(&free_bytes, &total_bytes));
        int64_t capacity = n * per_query;
        if (capacity > (int64_t)(free_bytes / 4 / entry_bytes)) capacity = (int64_t)(free_bytes / 4 / entry_bytes);
        if (capacity < 0) capacity = 0;
        CT_CHECK(count.alloc(1, s));
        CT_CUDA(cudaMemsetAsync(count.p, 0, sizeof(unsigned long long), s));
        CT_CHECK(q.alloc(capacity, s));
        CT_CHECK(k.alloc(capacity, s));
        CT_CHECK(j.alloc(capacity, s));
        if (with_xy) CT_CHECK(xy.alloc(4 * (size_t)capacity, s));
        log.count = count.p;
        log.capacity = capacity;
        log.q = q.p;
        log.k = k.p;
        log.j = j.p;
        log.xy = with_xy ? xy.p : nullptr;
        return CT_OK;
    }
};

// hits per query the log is sized for (ct_set_hit_log / CELLTREE_HIT_LOG; 0 = always traverse twice)
int64_t hit_log_per_query();

}  // namespace ct
