// edges.cu -- ct_intersect_edges: count -> scan -> fill traversal of query segments (Cohen-Sutherland +
// Cyrus-Beck against faces, segment/segment against a network), then the per-edge stable sort by t.
#include "hitlog.cuh"
#include "morton.cuh"
#include "traverse.cuh"

namespace ct {

// ---- edge kernels ----------------------------------------------------------------------------------------------
// The clip of a segment against a candidate cell (Cohen-Sutherland, then Cyrus-Beck) is by far the expensive part of
// the traversal, so it is done ONCE: count + log, scan, place (hitlog.cuh); the second traversal (FILL) only runs
// when the log overflows.
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
        counts[q] = locate_edge<MAXV>(t, a, b, [&](int k, int bbox_index, P2 c, P2 d) {
            const int64_t at = hitlog_reserve(log);
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
static int64_t g_hit_log = -1;
int64_t hit_log_per_query() {
    if (g_hit_log < 0) {
        const char *e = getenv("CELLTREE_HIT_LOG");
        g_hit_log = e ? atoll(e) : 16;
    }
    return g_hit_log;
}

template <int MAXV>
static int run_edges(const ct_tree *tree, const double *d_edges, int64_t n, ct_result *r, cudaStream_t s) {
    TreeView v = tree->view();
    Scratch<int32_t> counts;
    Scratch<int64_t> offsets;
    CT_CHECK(counts.alloc(n + 1, s));
    CT_CHECK(offsets.alloc(n + 1, s));
    CT_CUDA(cudaMemsetAsync(counts.p + n, 0, sizeof(int32_t), s));
    int64_t total = 0;
    MortonOrder order;
    CT_CHECK(order.build<KEY_EDGE>(tree, d_edges, n, s));
    HitLogBuffers buffers;
    CT_CHECK(buffers.alloc(n, hit_log_per_query(), true, s));
    const HitLog &log = buffers.log;
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
            k_place_hits<<<grid_for(total, 256), 256, 0, s>>>(log, total, offsets.p, r->i, r->j, r->payload);
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

extern "C" int ct_set_hit_log(int64_t hits_per_query) {
    g_hit_log = hits_per_query < 0 ? 16 : hits_per_query;
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

