// hitlog.cuh -- single-pass assembly of variable-length results.
//
// "Count, then fill" runs the traversal twice.  Where the per-candidate test is expensive (the segment clips of
// intersect_edges) the first pass instead counts the hits of every query AND keeps each hit: query q owns `per_query`
// slots, and its s-th hit (s = the value the per-query count had when the hit was counted: any order) goes to slot
// q * per_query + s with the ordinal of its candidate in the query's emission order, the cell and the payload.  The log is
// grouped by query from the start, so after the scan of the counts a warp sweeps the slots of 32 queries, ranks each query's
// hits and writes them to the query's range of the result, reading and writing whole lines (edges.cu: k_rank_slots).  A query with more
// hits than slots keeps its count but not the surplus hits; the caller redoes just those queries with the second traversal.
// (The first version appended all hits to one log in arrival order -- one global atomic per group of lanes -- and needed
// two kernels of scattered 4- to 40-byte accesses, 8.5 ms for the 86.5 M hits of C4, to bring them home.)
#pragma once

#include "common.cuh"

namespace ct {

constexpr int HIT_SLOTS_MAX = 32;  // k_rank_slots gives every slot of a query a lane

struct HitLog {
    int32_t per_query;  // slots per query; 0 = no log (two traversals)
    int2 *kj;           // (ordinal, cell) per slot
    double *xy;         // 4 doubles per slot
};

// Owns the log's device buffers for the duration of a call.
struct HitLogBuffers {
    Scratch<int2> kj;
    Scratch<double> xy;
    HitLog log{};

    // `per_query` slots per query, halved until the log fits a quarter of the free device memory (below 4: no log)
    int alloc(int64_t n, int64_t per_query, cudaStream_t s) {
        const int64_t entry_bytes = 40;
        size_t free_bytes = 0, total_bytes = 0;
        CT_CUDA(cudaMemGetInfo(&free_bytes, &total_bytes));
        if (per_query > HIT_SLOTS_MAX) per_query = HIT_SLOTS_MAX;
        while (per_query >= 4 && n * per_query > (int64_t)(free_bytes / 4 / entry_bytes)) per_query /= 2;
        if (per_query < 4 && per_query < hit_slots_requested()) per_query = 0;
        if (per_query < 0) per_query = 0;
        const int64_t capacity = n * per_query;
        CT_CHECK(kj.alloc(capacity, s));
        CT_CHECK(xy.alloc(4 * (size_t)capacity, s));
        log.per_query = (int32_t)per_query;
        log.kj = kj.p;
        log.xy = xy.p;
        return CT_OK;
    }
    static int64_t hit_slots_requested();
};

// slots per query the log is sized for (ct_set_hit_log / CELLTREE_HIT_LOG; 0 = always traverse twice)
int64_t hit_log_per_query();
inline int64_t HitLogBuffers::hit_slots_requested() { return hit_log_per_query(); }

}  // namespace ct
