// edges.cu -- ct_intersect_edges: count -> scan -> fill traversal of query segments (Cohen-Sutherland +
// Cyrus-Beck against faces, segment/segment against a network), then the per-edge stable sort by t.
#include "morton.cuh"
#include "traverse.cuh"

namespace ct {

// ---- edge kernels ----------------------------------------------------------------------------------------------
// The clip of a segment against a candidate cell (Cohen-Sutherland, then Cyrus-Beck) is by far the expensive part of
// the traversal, so it is done ONCE: the first pass counts the hits of every segment and appends each hit
// (segment, rank within the segment, cell, c, d) to a log, in whatever order the warps get there; after the scan of
// the counts a placement kernel moves every log entry to offsets[segment] + rank, which is the reference's order.
// If the log's capacity does not suffice the second traversal (FILL) writes the pairs instead.
struct HitLog {
    unsigned long long *count;  // entries requested so far (may exceed capacity)
    int64_t capacity;
    int32_t *q, *k, *j;
    double *xy;  // 4 doubles per entry
};

enum { EDGES_COUNT_AND_LOG = 0, EDGES_FILL = 1 };

template <int MAXV, int MODE>
__global__ void __launch_bounds__(BLOCK) k_locate_edges(TreeView t, const double *__restrict__ edges, int64_t n,
                                                        int32_t *__restrict__ counts, const int64_t *__restrict__ offsets,
                                                        int32_t *__restrict__ out_i, int32_t *__restrict__ out_j,
                                                        double *__restrict__ out_xy, const uint32_t *__restrict__ perm, HitLog log) {
    int64_t q = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
    if (q >= n) return;
    if (perm) q = __ldg(perm + q);  // execution order only
    const double2 *e = reinterpret_cast<const double2 *>(edges) + 2 * q;
    double2 a2 = __ldg(e), b2 = __ldg(e + 1);
    P2 a{a2.x, a2.y}, b{b2.x, b2.y};
    if constexpr (MODE == EDGES_FILL) {
        int64_t base = offsets[q];
        locate_edge<MAXV>(t, a, b, [&](int k, int bbox_index, P2 c, P2 d) {
            out_i[base + k] = (int32_t)q;
            out_j[base + k] = bbox_index;
            double2 *o = reinterpret_cast<double2 *>(out_xy + 4 * (base + k));
            o[0] = make_double2(c.x, c.y);
            o[1] = make_double2(d.x, d.y);
        });
    } else {
        const unsigned lane = threadIdx.x & 31u;
        counts[q] = locate_edge<MAXV>(t, a, b, [&](int k, int bbox_index, P2 c, P2 d) {
            // the lanes that arrive here together reserve their log entries with one atomic
            const unsigned mask = __activemask();
            const int leader = __ffs(mask) - 1;
            unsigned long long first = 0;
            if ((int)lane == leader) first = atomicAdd(log.count, (unsigned long long)__popc(mask));
            first = __shfl_sync(mask, first, leader);
            const int64_t at = (int64_t)first + __popc(mask & ((1u << lane) - 1u));
            if (at < log.capacity) {
                log.q[at] = (int32_t)q;
                log.k[at] = k;
                log.j[at] = bbox_index;
                double2 *o = reinterpret_cast<double2 *>(log.xy + 4 * at);
                o[0] = make_double2(c.x, c.y);
                o[1] = make_double2(d.x, d.y);
            }
        });
    }
}

// log entry -> its place in the result: offsets[segment] + rank within the segment
__global__ void __launch_bounds__(256) k_place_edges(HitLog log, int64_t entries, const int64_t *__restrict__ offsets,
                                                     int32_t *__restrict__ out_i, int32_t *__restrict__ out_j,
                                                     double *__restrict__ out_xy) {
    int64_t at = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (at >= entries) return;
    const int32_t q = __ldcs(log.q + at);
    const int64_t to = offsets[q] + __ldcs(log.k + at);
    out_i[to] = q;
    out_j[to] = __ldcs(log.j + at);
    const double2 *in = reinterpret_cast<const double2 *>(log.xy + 4 * at);
    double2 *o = reinterpret_cast<double2 *>(out_xy + 4 * to);
    o[0] = __ldcs(in);
    o[1] = __ldcs(in + 1);
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

}  // namespace ct

namespace ct {
static int64_t g_edge_log = -1;
int64_t edge_log_per_segment() {
    if (g_edge_log < 0) {
        const char *e = getenv("CELLTREE_EDGE_LOG");
        g_edge_log = e ? atoll(e) : 12;
    }
    return g_edge_log;
}

template <int MAXV>
static int run_edges(const ct_tree *tree, const double *d_edges, int64_t n, ct_result *r, cudaStream_t s) {
    TreeView v = tree->view();
    Scratch<int32_t> counts, log_q, log_k, log_j;
    Scratch<int64_t> offsets;
    Scratch<double> log_xy;
    Scratch<unsigned long long> log_count;
    CT_CHECK(counts.alloc(n + 1, s));
    CT_CHECK(offsets.alloc(n + 1, s));
    CT_CUDA(cudaMemsetAsync(counts.p + n, 0, sizeof(int32_t), s));
    int64_t total = 0;
    MortonOrder order;
    CT_CHECK(order.build<KEY_EDGE>(tree, d_edges, n, s));
    // room for 12 hits per segment on average (44 bytes each), at most a quarter of the free device memory
    HitLog log{};
    {
        const int64_t per_segment = edge_log_per_segment();
        size_t free_bytes = 0, total_bytes = 0;
        CT_CUDA(cudaMemGetInfo(&free_bytes, &total_bytes));
        int64_t capacity = n * per_segment;
        if (capacity > (int64_t)(free_bytes / 4 / 44)) capacity = (int64_t)(free_bytes / 4 / 44);
        if (capacity < 0) capacity = 0;
        CT_CHECK(log_count.alloc(1, s));
        CT_CUDA(cudaMemsetAsync(log_count.p, 0, sizeof(unsigned long long), s));
        CT_CHECK(log_q.alloc(capacity, s));
        CT_CHECK(log_k.alloc(capacity, s));
        CT_CHECK(log_j.alloc(capacity, s));
        CT_CHECK(log_xy.alloc(4 * (size_t)capacity, s));
        log.count = log_count.p;
        log.capacity = capacity;
        log.q = log_q.p;
        log.k = log_k.p;
        log.j = log_j.p;
        log.xy = log_xy.p;
    }
    if (n > 0) {
        k_locate_edges<MAXV, EDGES_COUNT_AND_LOG><<<grid_for(n, BLOCK), BLOCK, 0, s>>>(v, d_edges, n, counts.p, nullptr, nullptr, nullptr,
                                                                                      nullptr, order.perm, log);
        CT_LAUNCH_CHECK();
    }
    CT_CHECK(scan_counts(counts.p, n, offsets.p, &total, s));
    CT_CHECK(dalloc(&r->i, total, s));
    CT_CHECK(dalloc(&r->j, total, s));
    CT_CHECK(dalloc(&r->payload, total * 4, s));
    r->size = total;
    r->width = 4;
    if (n > 0 && total > 0) {
        if (total <= log.capacity) {
            k_place_edges<<<grid_for(total, 256), 256, 0, s>>>(log, total, offsets.p, r->i, r->j, r->payload);
        } else {
            k_locate_edges<MAXV, EDGES_FILL><<<grid_for(n, BLOCK), BLOCK, 0, s>>>(v, d_edges, n, nullptr, offsets.p, r->i, r->j,
                                                                                 r->payload, order.perm, log);
        }
        CT_LAUNCH_CHECK();
        k_sort_edge_ranges<<<grid_for(n, BLOCK), BLOCK, 0, s>>>(d_edges, n, offsets.p, r->j, r->payload);
        CT_LAUNCH_CHECK();
    }
    CT_CUDA(cudaStreamSynchronize(s));
    return CT_OK;
}
}  // namespace ct

using namespace ct;

extern "C" int ct_set_edge_log(int64_t hits_per_segment) {
    g_edge_log = hits_per_segment < 0 ? 12 : hits_per_segment;
    return CT_OK;
}

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

