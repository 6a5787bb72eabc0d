// edges.cu -- ct_intersect_edges: warp-cooperative traversal of the query segments that counts the hits and keeps them in
// per-segment slots (Cohen-Sutherland + Cyrus-Beck against faces, segment/segment against a network), scan, and the
// placement of every hit at its rank by (t, emission ordinal) within its segment's range.
#include "hitlog.cuh"
#include "morton.cuh"
#include "traverse.cuh"

namespace ct {

// ---- edge kernels ----------------------------------------------------------------------------------------------
// The clip of a segment against a candidate cell (Cohen-Sutherland, then Cyrus-Beck) is by far the expensive part, and
// with one thread per segment doing everything the lanes of a warp are hardly ever at the same place (ncu: 4.2 of 32
// lanes active per instruction).  The first pass is therefore WARP-COOPERATIVE:
//   * every lane walks the tree for its own segment, one node per iteration, through one branch-free step
//     (edge_plane_test + selects), and pushes the cells of the leaves it reaches -- with their ordinal in the
//     segment's candidate sequence -- onto a queue in shared memory that the warp shares;
//   * whenever the queue holds 32 candidates (or the walks are over), the 32 lanes take one candidate each, whoever
//     pushed it, and run the cheap half of the test (Cohen-Sutherland against the cell's box); the survivors go onto a
//     second queue, and whenever THAT holds 32 the lanes run the expensive half (Cyrus-Beck against the polygon) on one
//     survivor each; a hit is counted for its segment and kept in the slot of the segment's part of the log (hitlog.cuh)
//     that the count names, with the candidate's ordinal, the cell and the clipped points.
// After the scan of the counts, k_rank_slots gives every hit its rank by (t, ordinal) among its segment's hits and moves it
// from its slot to that place of the segment's range: the reference's stable sort by t of the hits in emission order
// (geometry_utils.py:564-574), since the ordinal grows with the emission order.  Handing the lanes of a warp further segments
// as they finish (a share of 64 .. 256 segments per warp) was measured and lost: C4 call 47.5 ms with 32 segments per warp,
// 49.0 / 49.9 / 52.9 ms with 64 / 128 / 256.  A segment with more hits than slots (32; on C4 a few dozen of 10 M) -- and
// every segment when there is no log -- takes the second traversal (one thread per segment, k_locate_edges_fill), which
// writes its pairs in emission order, and k_sort_edge_ranges sorts its range in place.
constexpr int QUEUE_CAP = 64;  // per warp: a step adds at most 32 candidates, a drain starts at 32

// 120 registers, 4 blocks of 128 threads per SM.  Capping the registers for more resident warps loses: C4 call 40.9 ms
// as is, 42.1 ms at 96 registers (5 blocks, 124 bytes spilled), 50.5 ms at 80 registers (6 blocks, 178 bytes spilled).
#ifndef CT_EDGES_MINB
#define CT_EDGES_MINB 4
#endif
template <int MAXV>
__global__ void __launch_bounds__(BLOCK, CT_EDGES_MINB) k_edges_cooperative(TreeView t, const double *__restrict__ edges, int64_t n,
                                                             int32_t *__restrict__ counts, const uint32_t *__restrict__ perm,
                                                             HitLog log) {
    constexpr int WARPS = BLOCK / 32;
    constexpr unsigned FULL = 0xffffffffu;
    __shared__ double2 s_segment[WARPS][32][2];
    __shared__ int32_t s_hits[WARPS][32];
    __shared__ int32_t s_cell[WARPS][QUEUE_CAP], s_ordinal[WARPS][QUEUE_CAP];    // candidates
    __shared__ uint8_t s_owner[WARPS][QUEUE_CAP];
    __shared__ int32_t s_cell2[WARPS][QUEUE_CAP], s_ordinal2[WARPS][QUEUE_CAP];  // candidates that passed the box clip
    __shared__ uint8_t s_owner2[WARPS][QUEUE_CAP];
    const int warp = threadIdx.x >> 5;
    const unsigned lane = threadIdx.x & 31u;
    const int64_t slot = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
    const bool valid = slot < n;
    const int64_t q = valid ? (perm ? (int64_t)__ldg(perm + slot) : slot) : 0;
    P2 a{0.0, 0.0}, b{0.0, 0.0};
    if (valid) {
        const double2 *e = reinterpret_cast<const double2 *>(edges) + 2 * q;
        const double2 a2 = __ldg(e), b2 = __ldg(e + 1);
        a = P2{a2.x, a2.y};
        b = P2{b2.x, b2.y};
    }
    s_segment[warp][lane][0] = make_double2(a.x, a.y);
    s_segment[warp][lane][1] = make_double2(b.x, b.y);
    s_hits[warp][lane] = 0;
    bool active = false;
    if (valid) {
        Box4 tree_bbox{t.bbox[0], t.bbox[1], t.bbox[2], t.bbox[3]};
        active = cohen_sutherland_line_meets_box(a, b, tree_bbox);
    }
    const P2 V = to_vector(a, b);
    const char *base = reinterpret_cast<const char *>(t.treelets);
    CT_STACK(stack, false);  // deep trees take the per-thread count / fill kernels instead (run_edges)
    Cursor cur;
    cursor_enter(cur, base, ROOT_HANDLE);
    int leaf_k = 0;    // cells of the current leaf already pushed
    int ordinal = 0;   // candidates of this segment pushed so far
    int queued = 0;    // entries in the warp's queues (warp-uniform)
    int queued2 = 0;
    __syncwarp();
    while (true) {
        const bool walking = __any_sync(FULL, active);
        if (!walking && queued == 0 && queued2 == 0) break;
        // ---- one node per walking lane ------------------------------------------------------------------------------
        bool push = false;
        int cell = 0;
        if (active) {
            uint32_t next = 0;
            bool pop = false, move = true;
            if (cursor_is_leaf(cur)) {
                const int4 leaf = cursor_leaf(cur);
                if (leaf_k < leaf.y) {
                    cell = leaf_element(leaf, t.bb_indices, leaf_k);
                    push = true;
                    leaf_k++;
                }
                if (leaf_k >= leaf.y) {
                    leaf_k = 0;
                    pop = true;
                } else {
                    move = false;  // more cells of this leaf to push
                }
            } else {
                bool left, right;
                edge_plane_test(cur, a, b, V, left, right);
                uint32_t left_handle, right_handle;
                cursor_children(cur, left_handle, right_handle);
                if (left && right) stack.push(t, left_handle);
                next = right ? right_handle : left_handle;
                pop = !(left || right);
            }
            if (pop) {
                if (stack.empty()) {
                    active = false;
                    move = false;
                } else {
                    next = stack.pop(t);
                }
            }
            if (move) cursor_enter(cur, base, next);
        }
        // ---- candidates onto the warp's queue -------------------------------------------------------------------------
        const unsigned pushing = __ballot_sync(FULL, push);
        if (push) {
            const int at = queued + __popc(pushing & ((1u << lane) - 1u));
            s_cell[warp][at] = cell;
            s_ordinal[warp][at] = ordinal++;
            s_owner[warp][at] = (uint8_t)lane;
        }
        queued += __popc(pushing);
        __syncwarp();
        // ---- 32 candidates at a time, one per lane ----------------------------------------------------------------
        const bool last = !__any_sync(FULL, active);
        // second stage: the expensive half on (up to) 32 survivors; `flush` also takes a partial batch
        auto clip_survivors = [&](bool flush) {
            while (queued2 >= 32 || (flush && queued2 > 0)) {
                const int take2 = queued2 < 32 ? queued2 : 32;
                const int first2 = queued2 - take2;
                const bool mine2 = (int)lane < take2;
                const int owner2 = mine2 ? s_owner2[warp][first2 + lane] : (int)lane;
                const int64_t owner_q = __shfl_sync(FULL, q, owner2);
                if (mine2) {
                    const int cell2 = s_cell2[warp][first2 + lane];
                    const double2 a2 = s_segment[warp][owner2][0], b2 = s_segment[warp][owner2][1];
                    P2 c, d;
                    bool hit;
                    if constexpr (MAXV == 0) hit = edge_edge_intersect(t, cell2, P2{a2.x, a2.y}, P2{b2.x, b2.y}, c, d);
                    else hit = edge_face_clip<MAXV>(t, cell2, P2{a2.x, a2.y}, P2{b2.x, b2.y}, c, d);
                    if (hit) {
                        // counted for its segment; the count it found is the hit's slot in the segment's part of the log
                        const int nth = atomicAdd(&s_hits[warp][owner2], 1);
                        if (nth < log.per_query) {
                            const int64_t at = owner_q * log.per_query + nth;
                            log.kj[at] = make_int2(s_ordinal2[warp][first2 + lane], cell2);
                            // the two points as ONE 256-bit store (sm_100: STG.E.ENL2.256): the slots of different segments
                            // are far apart, so every store instruction is a request of its own at the L2
                            asm volatile("st.global.v4.b64 [%0], {%1, %2, %3, %4};" ::"l"(log.xy + 4 * at),
                                         "l"(__double_as_longlong(c.x)), "l"(__double_as_longlong(c.y)),
                                         "l"(__double_as_longlong(d.x)), "l"(__double_as_longlong(d.y))
                                         : "memory");
                        }
                    }
                }
                queued2 = first2;
                __syncwarp();
            }
        };
        while (queued >= 32 || (last && queued > 0)) {
            const int take = queued < 32 ? queued : 32;
            const int first = queued - take;
            const bool mine = (int)lane < take;
            bool pass = false;
            int bbox_index = 0, owner = (int)lane, ordinal_of = 0;
            if (mine) {
                owner = s_owner[warp][first + lane];
                bbox_index = s_cell[warp][first + lane];
                ordinal_of = s_ordinal[warp][first + lane];
                const double2 a2 = s_segment[warp][owner][0], b2 = s_segment[warp][owner][1];
                if constexpr (MAXV == 0) pass = true;  // segment / segment: a single test, in the second stage
                else pass = edge_face_prefilter(t, bbox_index, P2{a2.x, a2.y}, P2{b2.x, b2.y});
            }
            queued = first;
            const unsigned passing = __ballot_sync(FULL, pass);
            if (pass) {
                const int at = queued2 + __popc(passing & ((1u << lane) - 1u));
                s_cell2[warp][at] = bbox_index;
                s_ordinal2[warp][at] = ordinal_of;
                s_owner2[warp][at] = (uint8_t)owner;
            }
            queued2 += __popc(passing);
            __syncwarp();
            clip_survivors(false);  // keeps the second queue below 32 entries
        }
        if (last) clip_survivors(true);  // the walks are over and the first queue is empty: flush
    }
    if (valid) counts[q] = s_hits[warp][lane];
}

// (x, ox) before (y, oy): smaller t first, NaN after every number, ties and NaNs by emission ordinal -- the order
// np.lexsort((t, edge)) gives the hits of one edge that arrive in emission order (geometry_utils.py:564-574)
CT_DEV bool hit_before(double x, int ox, double y, int oy) {
    if (x < y || (y != y && x == x)) return true;
    if (y < x || (x != x && y == y)) return false;
    return ox < oy;
}

// One warp per 32 consecutive segments, one segment (or two of at most 16 hits) per iteration, one LANE PER SLOT: the segment's logged hits are read in
// one coalesced sweep of its part of the log, every lane computes the key t = (c - a) . (b - a) of its hit, the keys go
// round the warp by shuffles, and a hit's rank among them under hit_before -- a strict total order, the ordinals of a
// segment's hits being distinct -- is its position in the segment's range of the result, where the lane writes it (the range is
// contiguous: the warp's stores fall into the same few lines).  A thread per segment walking its own slots was 8.0 ms on C4:
// every one of its accesses is a 32-byte request of its own, and the L2 takes those at 90 G/s.  Segments with more hits than
// slots are listed for the second traversal.
static_assert(HIT_SLOTS_MAX <= 32, "a lane per slot");
__global__ void __launch_bounds__(256) k_rank_slots(HitLog log, const double *__restrict__ edges, int64_t n,
                                                    const int32_t *__restrict__ counts, const int64_t *__restrict__ offsets,
                                                    int32_t *__restrict__ out_i, int32_t *__restrict__ out_j,
                                                    double *__restrict__ out_xy, int32_t *__restrict__ redo_count,
                                                    int32_t *__restrict__ redo_list) {
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int64_t q0 = (((int64_t)blockIdx.x * 256 + threadIdx.x) >> 5) << 5;
    if (q0 >= n) return;
    const int my_count = q0 + lane < n ? counts[q0 + lane] : 0;
    const int64_t my_lo = q0 + lane < n ? offsets[q0 + lane] : 0;
    if (my_count > log.per_query) redo_list[atomicAdd(redo_count, 1)] = (int32_t)(q0 + lane);
    const double2 *log_xy = reinterpret_cast<const double2 *>(log.xy);
    // (requesting the next segment's slots before the shuffles of the current one: 3.6 against 3.3 ms -- not latency)
    // `width` lanes per segment: 32, or 16 with segment i in the lower and segment i + 1 in the upper half of the warp
    auto run = [&](int i, int width) {
        const int half = width == 16 ? (lane >> 4) : 0;
        const int sub = width == 16 ? (lane & 15) : lane;
        const int k = __shfl_sync(FULL, my_count, i + half);
        const int64_t lo = __shfl_sync(FULL, my_lo, i + half);
        const bool ok = k > 0 && k <= log.per_query;
        const int rounds_mine = ok ? k : 0;
        const int rounds_lower = __shfl_sync(FULL, rounds_mine, 0), rounds_upper = __shfl_sync(FULL, rounds_mine, 16);
        const int rounds = rounds_lower > rounds_upper ? rounds_lower : rounds_upper;
        if (rounds == 0) return;  // the same for all lanes
        const int64_t q = q0 + i + half;
        const bool have = ok && sub < k;
        int2 kj = make_int2(0, 0);
        double2 a = make_double2(0.0, 0.0), b = a, c = a, d = a;
        if (have) {
            const double2 *e = reinterpret_cast<const double2 *>(edges) + 2 * q;
            a = __ldg(e);  // one address per segment
            b = __ldg(e + 1);
            const int64_t at = q * log.per_query + sub;
            kj = log.kj[at];
            c = log_xy[2 * at];
            d = log_xy[2 * at + 1];
        }
        const double abx = b.x - a.x, aby = b.y - a.y;
        const double key = (c.x - a.x) * abx + (c.y - a.y) * aby;
        const int first_lane = lane - sub;  // of this lane's segment
        int rank = 0;
        for (int s = 0; s < rounds; s++) {
            const int from = first_lane + (s < width ? s : 0);
            const double other = __shfl_sync(FULL, key, from);
            const int other_ordinal = __shfl_sync(FULL, kj.x, from);
            rank += (s < k && s != sub && hit_before(other, other_ordinal, key, kj.x)) ? 1 : 0;
        }
        if (have) {
            const int64_t to = lo + rank;
            out_i[to] = (int32_t)q;
            out_j[to] = kj.y;
            double2 *o = reinterpret_cast<double2 *>(out_xy + 4 * to);
            o[0] = c;
            o[1] = d;
        }
    };
    for (int i = 0; i < 32; i += 2) {
        const int k0 = __shfl_sync(FULL, my_count, i), k1 = __shfl_sync(FULL, my_count, i + 1);
        const bool small0 = k0 <= 16 || k0 > log.per_query, small1 = k1 <= 16 || k1 > log.per_query;  // (the latter are skipped)
        if (small0 && small1) {
            run(i, 16);  // 95 % of C4's segments have at most 16 hits: two segments per iteration
        } else {
            run(i, 32);
            run(i + 1, 32);
        }
    }
}

// the second traversal, one thread per segment: pairs in emission order (their ordinal is their rank)
// first pass for trees deeper than the per-thread stack: one thread per segment counts its hits (no log)
template <int MAXV>
__global__ void __launch_bounds__(BLOCK) k_locate_edges_count_deep(TreeView t, const double *__restrict__ edges, int64_t n,
                                                                   int32_t *__restrict__ counts, const uint32_t *__restrict__ perm) {
    int64_t q = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
    if (q >= n) return;
    if (perm) q = __ldg(perm + q);
    const double2 *e = reinterpret_cast<const double2 *>(edges) + 2 * q;
    double2 a2 = __ldg(e), b2 = __ldg(e + 1);
    counts[q] = locate_edge<MAXV, true>(t, P2{a2.x, a2.y}, P2{b2.x, b2.y}, [](int, int, P2, P2) {});
}

template <int MAXV, bool DEEP>
__global__ void __launch_bounds__(BLOCK) k_locate_edges_fill(TreeView t, const double *__restrict__ edges, int64_t n,
                                                             const int64_t *__restrict__ offsets, int32_t *__restrict__ out_i,
                                                             int32_t *__restrict__ out_j, double *__restrict__ out_xy,
                                                             int32_t *__restrict__ ordinal, const uint32_t *__restrict__ perm,
                                                             const int32_t *__restrict__ list) {
    int64_t q = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
    if (q >= n) return;  // n = entries of `list` (the segments whose hits did not fit the log) if there is one
    if (list) q = __ldg(list + q);
    else if (perm) q = __ldg(perm + q);  // execution order only
    const double2 *e = reinterpret_cast<const double2 *>(edges) + 2 * q;
    double2 a2 = __ldg(e), b2 = __ldg(e + 1);
    P2 a{a2.x, a2.y}, b{b2.x, b2.y};
    const int64_t base = offsets[q];
    locate_edge<MAXV, DEEP>(t, a, b, [&](int k, int bbox_index, P2 c, P2 d) {
        out_i[base + k] = (int32_t)q;
        out_j[base + k] = bbox_index;
        ordinal[base + k] = k;
        double2 *o = reinterpret_cast<double2 *>(out_xy + 4 * (base + k));
        o[0] = make_double2(c.x, c.y);
        o[1] = make_double2(d.x, d.y);
    });
}

// sort_intersections_by_edge, geometry_utils.py:564-574: within each query edge's (contiguous) range, the
// reference's np.lexsort((t, edge)) is a stable sort by t = (c - a) . (b - a) of the hits in emission order, NaN last.
// The range arrives in no particular order with the emission ordinal of every hit: sorting by (t, ordinal) is the same.
__global__ void __launch_bounds__(BLOCK) k_sort_edge_ranges(const double *__restrict__ edges, int64_t n,
                                                            const int64_t *__restrict__ offsets, int32_t *__restrict__ out_j,
                                                            double *__restrict__ out_xy, int32_t *__restrict__ ordinal,
                                                            const int32_t *__restrict__ list) {
    int64_t q = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
    if (q >= n) return;  // n = entries of `list` if there is one (the other segments came sorted out of the log)
    if (list) q = __ldg(list + q);
    int64_t lo = offsets[q], hi = offsets[q + 1];
    if (hi - lo < 2) return;
    const double2 *e = reinterpret_cast<const double2 *>(edges) + 2 * q;
    double2 a = __ldg(e), b = __ldg(e + 1);
    double abx = b.x - a.x, aby = b.y - a.y;
    double2 *xy = reinterpret_cast<double2 *>(out_xy);
    auto t_of = [&](double2 c) { return (c.x - a.x) * abx + (c.y - a.y) * aby; };
    for (int64_t k = lo + 1; k < hi; k++) {
        double2 c = xy[2 * k], d = xy[2 * k + 1];
        int32_t j = out_j[k], o = ordinal[k];
        double tk = t_of(c);
        int64_t m = k;
        while (m > lo) {
            double2 cp = xy[2 * (m - 1)];
            if (!hit_before(tk, o, t_of(cp), ordinal[m - 1])) break;
            xy[2 * m] = cp;
            xy[2 * m + 1] = xy[2 * (m - 1) + 1];
            out_j[m] = out_j[m - 1];
            ordinal[m] = ordinal[m - 1];
            m--;
        }
        if (m != k) {
            xy[2 * m] = c;
            xy[2 * m + 1] = d;
            out_j[m] = j;
            ordinal[m] = o;
        }
    }
}

}  // namespace ct

namespace ct {
static int64_t g_hit_log = -1;
int64_t hit_log_per_query() {
    if (g_hit_log < 0) {
        const char *e = getenv("CELLTREE_HIT_LOG");
        g_hit_log = e ? atoll(e) : HIT_SLOTS_MAX;
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
    DeepScope deep;
    CT_CHECK(deep.init(tree, n, s));
    v.deep = deep.view;
    const bool deep_tree = deep.view.slab != nullptr;
    HitLogBuffers buffers;
    CT_CHECK(buffers.alloc(n, deep_tree ? 0 : hit_log_per_query(), s));
    const HitLog &log = buffers.log;
    if (n > 0) {
        CT_CHECK(deep.next_launch());
        if (deep_tree) k_locate_edges_count_deep<MAXV><<<grid_for(n, BLOCK), BLOCK, 0, s>>>(v, d_edges, n, counts.p, order.perm);
        else k_edges_cooperative<MAXV><<<grid_for(n, BLOCK), BLOCK, 0, s>>>(v, d_edges, n, counts.p, order.perm, log);
        CT_LAUNCH_CHECK();
    }
    CT_CHECK(scan_counts(counts.p, n, offsets.p, &total, s));
    CT_CHECK(dalloc(&r->i, total, s));
    CT_CHECK(dalloc(&r->j, total, s));
    CT_CHECK(dalloc(&r->payload, total * 4, s));
    r->size = total;
    r->width = 4;
    if (n > 0 && total > 0) {
        // the segments whose hits are all in the log are ranked and written from there; the others (all of them without a
        // log) are listed and take the second traversal: pairs in emission order, then the per-segment insertion sort
        int32_t redo = -1;  // -1: every segment
        Scratch<int32_t> redo_list, redo_count;
        if (log.per_query > 0) {
            CT_CHECK(redo_list.alloc(n, s));
            CT_CHECK(redo_count.alloc(1, s));
            CT_CUDA(cudaMemsetAsync(redo_count.p, 0, sizeof(int32_t), s));
            k_rank_slots<<<grid_for(n, 256), 256, 0, s>>>(log, d_edges, n, counts.p, offsets.p, r->i, r->j, r->payload, redo_count.p,
                                                          redo_list.p);
            CT_LAUNCH_CHECK();
            CT_CUDA(cudaMemcpyAsync(&redo, redo_count.p, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
            CT_CUDA(cudaStreamSynchronize(s));
        }
        if (redo != 0) {
            const int64_t n_redo = redo < 0 ? n : redo;
            const int32_t *list = redo < 0 ? nullptr : redo_list.p;
            Scratch<int32_t> ordinal;
            CT_CHECK(ordinal.alloc(total, s));
            CT_CHECK(deep.next_launch());
            if (deep_tree)
                k_locate_edges_fill<MAXV, true><<<grid_for(n_redo, BLOCK), BLOCK, 0, s>>>(v, d_edges, n_redo, offsets.p, r->i, r->j,
                                                                                         r->payload, ordinal.p, order.perm, list);
            else
                k_locate_edges_fill<MAXV, false><<<grid_for(n_redo, BLOCK), BLOCK, 0, s>>>(v, d_edges, n_redo, offsets.p, r->i, r->j,
                                                                                          r->payload, ordinal.p, order.perm, list);
            CT_LAUNCH_CHECK();
            k_sort_edge_ranges<<<grid_for(n_redo, BLOCK), BLOCK, 0, s>>>(d_edges, n_redo, offsets.p, r->j, r->payload, ordinal.p, list);
            CT_LAUNCH_CHECK();
        }
    }
    CT_CUDA(cudaStreamSynchronize(s));
    return deep.finish();
}
}  // namespace ct

using namespace ct;

extern "C" int ct_set_hit_log(int64_t hits_per_query) {
    g_hit_log = hits_per_query < 0 ? HIT_SLOTS_MAX : hits_per_query;
    return CT_OK;
}

extern "C" int ct_intersect_edges(const ct_tree *tree, const double *edges, int64_t n, int32_t mem, ct_result **out) {
    if (!tree || !out || n < 0 || (n > 0 && !edges)) {
        set_error("ct_intersect_edges: null argument");
        return CT_ERR_VALUE;
    }
    CT_ON_DEVICE(tree->device);
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

