// build.cu -- mesh preparation and level-synchronous cell tree construction on the GPU.
//
// Replaces the serial constructor pipeline of the reference (celltree.py:74-97):
//   counter_clockwise (geometry_utils.py:541-561), build_face_bboxes / build_edge_bboxes (:443-487),
//   creation.initialize / build (creation.py:233-413), bbox_tree / bbox_distances / default_tolerance
//   (celltree_base.py:20-52).
//
// The reference builds depth-first with an explicit stack; every node only permutes its own slice
// [ptr, ptr + size) of bb_indices, so the result is independent of the processing order EXCEPT for the node
// numbering.  Here all nodes of one "wave" are processed together, element-parallel:
//
//   k_range    per element: min of box-min / max of box-max of its node in the node's dim   (get_bounds, :153-171)
//   k_bucket   per element: bucket of the bbox centroid (first bucket whose [Min, Max) holds it, :44-51, :278-293)
//              + per (node, bucket) count, Rmin, Lmax                                      (:301-306)
//   k_decide   per node: special case cells_per_leaf == 1 (:312-320), drop empty buckets (:322-341), retry in the
//              other dim / oversized leaf (:345-358), split_plane (:174-213), create the two children (:365-379)
//   scan+k_rank+k_scatter  stable partition of each node's slice by bucket (sort_bbox_indices / stable_partition,
//              :54-150): new position = ptr + (elements in lower buckets) + (rank among same-bucket elements)
//
// and afterwards the nodes are renumbered to the reference's numbering: the k-th SPLITTING node in left-first
// pre-order owns children 1 + 2k and 2 + 2k (:373-379).  Subtree split counts bottom-up, pre-order ranks
// top-down, one kernel per wave each.
//
// min / max over doubles use atomicMin / atomicMax on an order-preserving uint64 image of the double.
#include <cub/device/device_scan.cuh>
#include <cub/iterator/counting_input_iterator.cuh>
#include <cub/iterator/transform_input_iterator.cuh>

#include <cstring>
#include <ctime>
#include <vector>

#include "traverse.cuh"

namespace ct {

constexpr int BB = 256;

// ---- ordered image of a double ---------------------------------------------------------------------------
CT_DEV unsigned long long enc(double d) {
    unsigned long long u = (unsigned long long)__double_as_longlong(d);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ULL);
}
CT_DEV double dec(unsigned long long u) {
    u = (u >> 63) ? (u & 0x7fffffffffffffffULL) : ~u;
    return __longlong_as_double((long long)u);
}
static inline unsigned long long enc_host(double d) {
    unsigned long long u;
    memcpy(&u, &d, 8);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ULL);
}
static inline double dec_host(unsigned long long u) {
    u = (u >> 63) ? (u & 0x7fffffffffffffffULL) : ~u;
    double d;
    memcpy(&d, &u, 8);
    return d;
}

// ---- small conversion kernels ----------------------------------------------------------------------------
__global__ void __launch_bounds__(BB) k_widen(const int32_t *__restrict__ in, int64_t n, int64_t *__restrict__ out) {
    int64_t k = (int64_t)blockIdx.x * BB + threadIdx.x;
    if (k < n) out[k] = in[k];
}
__global__ void __launch_bounds__(BB) k_narrow(const int64_t *__restrict__ in, int64_t n, int32_t *__restrict__ out, int64_t fill) {
    int64_t k = (int64_t)blockIdx.x * BB + threadIdx.x;
    if (k < n) {
        const int64_t v = in[k];
        out[k] = v == fill ? -1 : (int32_t)v;
    }
}
int launch_widen(const int32_t *in, int64_t n, int64_t *out, cudaStream_t s) {
    if (n <= 0) return CT_OK;
    k_widen<<<grid_for(n, BB), BB, 0, s>>>(in, n, out);
    CT_LAUNCH_CHECK();
    return CT_OK;
}
int launch_narrow(const int64_t *in, int64_t n, int32_t *out, cudaStream_t s, int64_t fill) {
    if (n <= 0) return CT_OK;
    k_narrow<<<grid_for(n, BB), BB, 0, s>>>(in, n, out, fill);
    CT_LAUNCH_CHECK();
    return CT_OK;
}

// ---- counter_clockwise, geometry_utils.py:541-561 ----------------------------------------------------------
// Literal restatement: a stateful loop, NOT a signed-area test -- after a flip the loop carries on with the
// stale a, b, exactly as the reference does.
__global__ void __launch_bounds__(BB) k_counter_clockwise(const double2 *__restrict__ vertices, int32_t *__restrict__ faces,
                                                          int64_t n_face, int M) {
    int64_t f = (int64_t)blockIdx.x * BB + threadIdx.x;
    if (f >= n_face) return;
    int32_t *face = faces + f * M;
    int length = M;  // polygon_length, :72-79
    for (int i = 3; i < M; i++)
        if (face[i] == -1) {
            length = i;
            break;
        }
    double2 a2 = vertices[face[length - 2]], b2 = vertices[face[length - 1]];
    P2 a{a2.x, a2.y}, b{b2.x, b2.y};
    for (int i = 0; i < length; i++) {
        double2 c2 = vertices[face[i]];
        P2 c{c2.x, c2.y};
        P2 u = to_vector(a, b);
        P2 v = to_vector(a, c);
        double product = cross_product(u, v);
        if (product == 0) {
            a = b;
            b = c;
        } else if (product < 0) {
            int end = length - 1;  // flip, :532-538
            for (int k = 0; k < length / 2; k++) {
                int32_t t = face[k];
                face[k] = face[end - k];
                face[end - k] = t;
            }
        } else {
            break;
        }
    }
}
int launch_counter_clockwise(const double2 *vertices, int32_t *faces, int64_t n_face, int M, cudaStream_t s) {
    if (n_face <= 0) return CT_OK;
    k_counter_clockwise<<<grid_for(n_face, BB), BB, 0, s>>>(vertices, faces, n_face, M);
    CT_LAUNCH_CHECK();
    return CT_OK;
}

// ---- build_face_bboxes / bounding_box, geometry_utils.py:421-456 ----------------------------------------
__global__ void __launch_bounds__(BB) k_face_bboxes(const double2 *__restrict__ vertices, const int32_t *__restrict__ faces,
                                                    int64_t n_face, int M, double *__restrict__ bb) {
    int64_t f = (int64_t)blockIdx.x * BB + threadIdx.x;
    if (f >= n_face) return;
    const int32_t *face = faces + f * M;
    double2 first = vertices[face[0]];
    double xmin = first.x, xmax = first.x, ymin = first.y, ymax = first.y;
    for (int k = 1; k < M; k++) {
        int index = face[k];
        if (index == -1) break;
        double2 v = vertices[index];
        xmin = nb_min(xmin, v.x);
        xmax = nb_max(xmax, v.x);
        ymin = nb_min(ymin, v.y);
        ymax = nb_max(ymax, v.y);
    }
    double2 *o = reinterpret_cast<double2 *>(bb + 4 * f);
    o[0] = make_double2(xmin, xmax);
    o[1] = make_double2(ymin, ymax);
}
int launch_face_bboxes(const double2 *vertices, const int32_t *faces, int64_t n_face, int M, double *bb, cudaStream_t s) {
    if (n_face <= 0) return CT_OK;
    k_face_bboxes<<<grid_for(n_face, BB), BB, 0, s>>>(vertices, faces, n_face, M, bb);
    CT_LAUNCH_CHECK();
    return CT_OK;
}

// ---- build_edge_bboxes / edge_bounding_box, geometry_utils.py:459-487 ----------------------------------
__global__ void __launch_bounds__(BB) k_edge_bboxes(const double2 *__restrict__ vertices, const int32_t *__restrict__ edges,
                                                    int64_t n_edge, double tolerance, double *__restrict__ bb) {
    int64_t e = (int64_t)blockIdx.x * BB + threadIdx.x;
    if (e >= n_edge) return;
    double2 v0 = vertices[edges[2 * e]], v1 = vertices[edges[2 * e + 1]];
    double2 *o = reinterpret_cast<double2 *>(bb + 4 * e);
    o[0] = make_double2(nb_min(v0.x - tolerance, v1.x - tolerance), nb_max(v0.x + tolerance, v1.x + tolerance));
    o[1] = make_double2(nb_min(v0.y - tolerance, v1.y - tolerance), nb_max(v0.y + tolerance, v1.y + tolerance));
}

// ---- bbox_tree + max bbox diagonal, celltree_base.py:20-52 --------------------------------------------
// red[0..3] = enc(min xmin), enc(max xmax), enc(min ymin), enc(max ymax); red[4] = enc(max diagonal)
__global__ void __launch_bounds__(BB) k_bbox_reduce(const double *__restrict__ bb, int64_t n, unsigned long long *__restrict__ red) {
    double xmin = FLOAT_MAX, xmax = FLOAT_MIN, ymin = FLOAT_MAX, ymax = FLOAT_MIN, diag = FLOAT_MIN;
    for (int64_t i = (int64_t)blockIdx.x * BB + threadIdx.x; i < n; i += (int64_t)gridDim.x * BB) {
        Box4 b = load_box(bb, i);
        xmin = b.xmin < xmin ? b.xmin : xmin;
        xmax = b.xmax > xmax ? b.xmax : xmax;
        ymin = b.ymin < ymin ? b.ymin : ymin;
        ymax = b.ymax > ymax ? b.ymax : ymax;
        double dx = b.xmax - b.xmin, dy = b.ymax - b.ymin;
        double d = sqrt(dx * dx + dy * dy);  // bbox_distances, celltree_base.py:41-47
        diag = d > diag ? d : diag;
    }
    unsigned long long e0 = enc(xmin), e1 = enc(xmax), e2 = enc(ymin), e3 = enc(ymax), e4 = enc(diag);
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long t;
        t = __shfl_xor_sync(0xffffffffu, e0, o); e0 = t < e0 ? t : e0;
        t = __shfl_xor_sync(0xffffffffu, e1, o); e1 = t > e1 ? t : e1;
        t = __shfl_xor_sync(0xffffffffu, e2, o); e2 = t < e2 ? t : e2;
        t = __shfl_xor_sync(0xffffffffu, e3, o); e3 = t > e3 ? t : e3;
        t = __shfl_xor_sync(0xffffffffu, e4, o); e4 = t > e4 ? t : e4;
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(red + 0, e0);
        atomicMax(red + 1, e1);
        atomicMin(red + 2, e2);
        atomicMax(red + 3, e3);
        atomicMax(red + 4, e4);
    }
}

__global__ void __launch_bounds__(BB) k_bb_distances(const double *__restrict__ bb, int64_t n, double *__restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * BB + threadIdx.x;
    if (i >= n) return;
    Box4 b = load_box(bb, i);
    double dx = b.xmax - b.xmin, dy = b.ymax - b.ymin;
    out[3 * i] = dx;
    out[3 * i + 1] = dy;
    out[3 * i + 2] = sqrt(dx * dx + dy * dy);
}

// ---- NodeDType (41 bytes, packed) <-> Node32 -----------------------------------------------------------------
// A block converts 256 nodes; the packed side is staged through shared memory so that global traffic is
// coalesced 4-byte words (256 * 41 bytes is a multiple of 4).
constexpr int NODE41 = 41;

__global__ void __launch_bounds__(BB) k_pack_nodes(const Node32 *__restrict__ nodes, int64_t n, unsigned char *__restrict__ out) {
    __shared__ __align__(16) unsigned char buf[BB * NODE41];
    int64_t base = (int64_t)blockIdx.x * BB;
    int64_t i = base + threadIdx.x;
    if (i < n) {
        Node32 nd = nodes[i];
        ct_node41 p;
        p.child = nd.child;
        p.Lmax = nd.Lmax;
        p.Rmin = nd.Rmin;
        p.ptr = nd.ptr;
        p.size = nd.size;
        p.dim = (uint8_t)(nd.dim ? 1 : 0);
        memcpy(buf + threadIdx.x * NODE41, &p, NODE41);
    }
    __syncthreads();
    int64_t count = (n - base) < BB ? (n - base) : BB;
    int64_t bytes = count * NODE41;
    unsigned char *dst = out + base * NODE41;  // base * 41 is a multiple of 4
    int64_t words = bytes / 4;
    for (int64_t w = threadIdx.x; w < words; w += BB) reinterpret_cast<uint32_t *>(dst)[w] = reinterpret_cast<uint32_t *>(buf)[w];
    for (int64_t b = words * 4 + threadIdx.x; b < bytes; b += BB) dst[b] = buf[b];
}

__global__ void __launch_bounds__(BB) k_unpack_nodes(const unsigned char *__restrict__ in, int64_t n, Node32 *__restrict__ nodes) {
    __shared__ __align__(16) unsigned char buf[BB * NODE41];
    int64_t base = (int64_t)blockIdx.x * BB;
    int64_t count = (n - base) < BB ? (n - base) : BB;
    int64_t bytes = count * NODE41;
    const unsigned char *src = in + base * NODE41;
    int64_t words = bytes / 4;
    for (int64_t w = threadIdx.x; w < words; w += BB) reinterpret_cast<uint32_t *>(buf)[w] = reinterpret_cast<const uint32_t *>(src)[w];
    for (int64_t b = words * 4 + threadIdx.x; b < bytes; b += BB) buf[b] = src[b];
    __syncthreads();
    int64_t i = base + threadIdx.x;
    if (i < n) {
        ct_node41 p;
        memcpy(&p, buf + threadIdx.x * NODE41, NODE41);
        Node32 nd;
        nd.child = (int32_t)p.child;
        nd.Lmax = p.Lmax;
        nd.Rmin = p.Rmin;
        nd.ptr = (int32_t)p.ptr;
        nd.size = (int32_t)p.size;
        nd.dim = p.dim ? 1 : 0;
        nodes[i] = nd;
    }
}

// depth of an uploaded tree: parent links, then every node counts its ancestors
__global__ void __launch_bounds__(BB) k_parents(const Node32 *__restrict__ nodes, int64_t n, int32_t *__restrict__ parent) {
    int64_t i = (int64_t)blockIdx.x * BB + threadIdx.x;
    if (i >= n) return;
    int c = nodes[i].child;
    if (c >= 0 && c + 1 < n) {
        parent[c] = (int32_t)i;
        parent[c + 1] = (int32_t)i;
    }
}
__global__ void __launch_bounds__(BB) k_depth(const int32_t *__restrict__ parent, int64_t n, int32_t *__restrict__ max_depth) {
    int64_t i = (int64_t)blockIdx.x * BB + threadIdx.x;
    if (i >= n) return;
    int d = 1;
    int p = parent[i];
    while (p >= 0 && d < (1 << 20)) {
        d++;
        p = parent[p];
    }
    atomicMax(max_depth, d);
}

// ==============================================================================================================
// Level-synchronous build
// ==============================================================================================================
struct BuildState {
    // elements
    const double *bb;
    int32_t *seg;  // slot of the active node that owns this position, -1 when the position is final
    uint16_t *bkt;  // bucket of the element at this position (valid where seg >= 0); n_buckets <= 65535
    // nodes in creation order
    int32_t *n_ptr, *n_size, *n_child;
    uint8_t *n_dim, *n_tried;
    double *n_Lmax, *n_Rmin;
    // per active slot
    const int32_t *active;
    unsigned long long *a_min, *a_max;
    int32_t *b_cnt;                    // [slot * nb + k]
    unsigned long long *b_min, *b_max;  // [slot * nb + k]
    int32_t *b_start;                  // [slot * nb + k] exclusive prefix of b_cnt
    int32_t *next_left, *next_right, *split_pos;
    int32_t *next_active;
    int32_t *counters;  // [0] node count, [1] next active count, [2] error flag
    int nb, cpl;
    // Signed-zero mode (only when a bounding box holds -0.0, see k_zero_first): position of the first element whose
    // value equals a zero extreme, per active slot (range) and per (slot, bucket); nullptr otherwise
    int32_t *a_minpos, *a_maxpos, *b_minpos, *b_maxpos;
    // At most four buckets (the default): the stable partition's ranks come from per-block bucket counts (written by
    // k_bucket) instead of a scan over all positions, see k_scatter4.  nullptr otherwise.
    uint4 *block_counts;
    uint4 *node_inblock;  // per active slot: bucket counts within its first position's block, before that position
};

// The reference takes extremes with strict comparisons in slice order (get_bounds, creation.py:153-171: `value < Rmin`,
// `value > Lmax`), so among -0.0 and +0.0 -- equal as numbers, different as bits -- the one met FIRST is kept, and that
// is what ends up in nodes["Lmax"] / nodes["Rmin"].  The ordered integer image separates the two zeros (-0.0 below
// +0.0), so atomic min / max would always keep -0.0 / +0.0.  When the boxes hold a -0.0 (a flag computed once per
// build; otherwise none of this runs), zeros enter the atomics as +0.0, and wherever an extreme comes out as zero the
// element that supplies it is found by position: k_zero_first takes the minimum position among the elements whose
// value is a zero, k_zero_apply reads that element's own zero back.
CT_DEV double unsigned_zero(double v) { return v == 0.0 ? 0.0 : v; }

__global__ void __launch_bounds__(BB) k_init_slots(BuildState st, int n_active) {
    int s = blockIdx.x * BB + threadIdx.x;
    if (s >= n_active) return;
    st.a_min[s] = enc(FLOAT_MAX);   // get_bounds starts from FLOAT_MAX / FLOAT_MIN, creation.py:161-162
    st.a_max[s] = enc(FLOAT_MIN);
    if (st.a_minpos) {
        st.a_minpos[s] = INT32_MAX;
        st.a_maxpos[s] = INT32_MAX;
    }
    for (int k = 0; k < st.nb; k++) {
        st.b_cnt[(int64_t)s * st.nb + k] = 0;
        st.b_min[(int64_t)s * st.nb + k] = enc(FLOAT_MAX);
        st.b_max[(int64_t)s * st.nb + k] = enc(FLOAT_MIN);
        if (st.a_minpos) {
            st.b_minpos[(int64_t)s * st.nb + k] = INT32_MAX;
            st.b_maxpos[(int64_t)s * st.nb + k] = INT32_MAX;
        }
    }
}

// get_bounds over a node's slice (creation.py:153-171).  `value < Rmin` / `value > Lmax` never pick a NaN.
__global__ void __launch_bounds__(BB) k_range(BuildState st, const int32_t *__restrict__ idx, int64_t n) {
    __shared__ unsigned long long s_min[BB / 32], s_max[BB / 32];
    int64_t pos = (int64_t)blockIdx.x * BB + threadIdx.x;
    int64_t first = (int64_t)blockIdx.x * BB;
    int64_t last = first + BB - 1 < n - 1 ? first + BB - 1 : n - 1;
    int slot_first = st.seg[first], slot_last = st.seg[last];
    // a node's positions are contiguous: equal slots at both ends of the block => one node owns the block
    bool uniform = (slot_first == slot_last) && slot_first >= 0;
    int slot = pos < n ? st.seg[pos] : -1;
    unsigned long long emin = enc(FLOAT_MAX), emax = enc(FLOAT_MIN);
    if (slot >= 0) {
        int node = st.active[slot];
        int dim = st.n_dim[node];
        int e = idx[pos];
        double vmin = st.bb[4 * (int64_t)e + 2 * dim];
        double vmax = st.bb[4 * (int64_t)e + 2 * dim + 1];
        if (st.a_minpos) vmin = unsigned_zero(vmin), vmax = unsigned_zero(vmax);
        if (vmin == vmin) emin = enc(vmin);
        if (vmax == vmax) emax = enc(vmax);
    }
    if (!uniform) {
        if (slot >= 0) {
            atomicMin(st.a_min + slot, emin);
            atomicMax(st.a_max + slot, emax);
        }
        return;
    }
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long t;
        t = __shfl_xor_sync(0xffffffffu, emin, o); emin = t < emin ? t : emin;
        t = __shfl_xor_sync(0xffffffffu, emax, o); emax = t > emax ? t : emax;
    }
    if ((threadIdx.x & 31) == 0) {
        s_min[threadIdx.x >> 5] = emin;
        s_max[threadIdx.x >> 5] = emax;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < BB / 32; w++) {
            emin = s_min[w] < emin ? s_min[w] : emin;
            emax = s_max[w] > emax ? s_max[w] : emax;
        }
        atomicMin(st.a_min + slot_first, emin);
        atomicMax(st.a_max + slot_first, emax);
    }
}

// bucket-k elements among the earlier threads of the block, k = 0..3 (used by the at-most-four-buckets partition below)
__device__ __forceinline__ uint4 block_exclusive4(int bucket, uint4 (*s_warp)[BB / 32]) {
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    const unsigned m0 = __ballot_sync(0xffffffffu, bucket == 0), m1 = __ballot_sync(0xffffffffu, bucket == 1);
    const unsigned m2 = __ballot_sync(0xffffffffu, bucket == 2), m3 = __ballot_sync(0xffffffffu, bucket == 3);
    if (lane == 0) (*s_warp)[warp] = make_uint4(__popc(m0), __popc(m1), __popc(m2), __popc(m3));
    __syncthreads();
    uint4 r = make_uint4(__popc(m0 & lt), __popc(m1 & lt), __popc(m2 & lt), __popc(m3 & lt));
    for (unsigned w = 0; w < warp; w++) {
        const uint4 t = (*s_warp)[w];
        r.x += t.x, r.y += t.y, r.z += t.z, r.w += t.w;
    }
    __syncthreads();
    return r;
}

constexpr int SMEM_BUCKETS = 64;

// Bucket of every element (centroid_test, creation.py:44-51, applied bucket after bucket by
// sort_bbox_indices :115-150: the FIRST bucket whose half-open range holds the centroid) and per-bucket
// count / Rmin / Lmax (get_bounds per bucket, :301-306).
__global__ void __launch_bounds__(BB) k_bucket(BuildState st, const int32_t *__restrict__ idx, int64_t n) {
    __shared__ int s_cnt[SMEM_BUCKETS];
    __shared__ unsigned long long s_bmin[SMEM_BUCKETS], s_bmax[SMEM_BUCKETS];
    const int nb = st.nb;
    int64_t pos = (int64_t)blockIdx.x * BB + threadIdx.x;
    int64_t first = (int64_t)blockIdx.x * BB;
    int64_t last = first + BB - 1 < n - 1 ? first + BB - 1 : n - 1;
    int slot_first = st.seg[first], slot_last = st.seg[last];
    bool uniform = (slot_first == slot_last) && slot_first >= 0 && nb <= SMEM_BUCKETS;
    if (uniform) {
        for (int k = threadIdx.x; k < nb; k += BB) {
            s_cnt[k] = 0;
            s_bmin[k] = enc(FLOAT_MAX);
            s_bmax[k] = enc(FLOAT_MIN);
        }
        __syncthreads();
    }
    int slot = pos < n ? st.seg[pos] : -1;
    int my_bucket = -1;
    if (slot >= 0) {
        int node = st.active[slot];
        int dim = st.n_dim[node];
        int e = idx[pos];
        double vmin = st.bb[4 * (int64_t)e + 2 * dim];
        double vmax = st.bb[4 * (int64_t)e + 2 * dim + 1];
        double range_Rmin = dec(st.a_min[slot]);
        double range_Lmax = dec(st.a_max[slot]);
        double bucket_length = (range_Lmax - range_Rmin) / (double)nb;  // creation.py:278
        double centroid = vmin + 0.5 * (vmax - vmin);                    // creation.py:50
        int k = -1;
        if (nb > 32 && bucket_length > 0.0 && bucket_length <= FLOAT_MAX) {
            // Many buckets: b -> (double)b * bucket_length + range_Rmin is non-decreasing for a finite positive length
            // (rounding is monotone), and bmax of bucket b is computed exactly like bmin of bucket b + 1.  So
            // "centroid < bmax(b)" holds from some first bucket on, "centroid >= bmin(b)" up to some last one, and the
            // FIRST bucket satisfying both -- what the loop below finds -- is the first one with centroid < bmax(b),
            // if its bmin test holds too.  A NaN centroid fails every test in both forms.
            int lo = 0, hi = nb;  // first b in [0, nb) with centroid < bmax(b); nb if none
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                const double bmax = (double)(mid + 1) * bucket_length + range_Rmin;
                if (centroid < bmax) hi = mid;
                else lo = mid + 1;
            }
            if (lo < nb && centroid >= (double)lo * bucket_length + range_Rmin) k = lo;
        } else {
            for (int b = 0; b < nb; b++) {
                double bmax = (double)(b + 1) * bucket_length + range_Rmin;  // creation.py:286
                double bmin = (double)b * bucket_length + range_Rmin;        // creation.py:287
                if ((centroid >= bmin) && (centroid < bmax)) {
                    k = b;
                    break;
                }
            }
        }
        if (k < 0) {
            atomicExch(st.counters + 2, CT_ERR_UNBUCKETABLE);
            k = nb - 1;
        }
        st.bkt[pos] = (uint16_t)k;
        my_bucket = k;
        if (st.a_minpos) vmin = unsigned_zero(vmin), vmax = unsigned_zero(vmax);
        unsigned long long emin = (vmin == vmin) ? enc(vmin) : enc(FLOAT_MAX);
        unsigned long long emax = (vmax == vmax) ? enc(vmax) : enc(FLOAT_MIN);
        if (uniform) {
            atomicAdd(&s_cnt[k], 1);
            atomicMin(&s_bmin[k], emin);
            atomicMax(&s_bmax[k], emax);
        } else {
            int64_t o = (int64_t)slot * nb + k;
            atomicAdd(st.b_cnt + o, 1);
            atomicMin(st.b_min + o, emin);
            atomicMax(st.b_max + o, emax);
        }
    }
    if (uniform) {
        __syncthreads();
        for (int k = threadIdx.x; k < nb; k += BB) {
            if (s_cnt[k] > 0) {
                int64_t o = (int64_t)slot_first * nb + k;
                atomicAdd(st.b_cnt + o, s_cnt[k]);
                atomicMin(st.b_min + o, s_bmin[k]);
                atomicMax(st.b_max + o, s_bmax[k]);
            }
        }
    }
    if (st.block_counts) {  // how many elements of this block went to each of the (at most four) buckets
        const int c0 = __syncthreads_count(my_bucket == 0), c1 = __syncthreads_count(my_bucket == 1);
        const int c2 = __syncthreads_count(my_bucket == 2), c3 = __syncthreads_count(my_bucket == 3);
        if (threadIdx.x == 0) st.block_counts[blockIdx.x] = make_uint4(c0, c1, c2, c3);
        __shared__ uint4 s_warp4[BB / 32];
        const uint4 before = block_exclusive4(my_bucket, &s_warp4);
        if (slot >= 0 && pos == st.n_ptr[st.active[slot]]) st.node_inblock[slot] = before;
    }
}

// Signed-zero mode, after k_bucket: first position per slot / bucket whose value is a zero, where the extreme is zero.
__global__ void __launch_bounds__(BB) k_zero_first(BuildState st, const int32_t *__restrict__ idx, int64_t n) {
    const int64_t pos = (int64_t)blockIdx.x * BB + threadIdx.x;
    if (pos >= n) return;
    const int slot = st.seg[pos];
    if (slot < 0) return;
    const int dim = st.n_dim[st.active[slot]];
    const int e = idx[pos];
    const double vmin = st.bb[4 * (int64_t)e + 2 * dim], vmax = st.bb[4 * (int64_t)e + 2 * dim + 1];
    const int64_t o = (int64_t)slot * st.nb + st.bkt[pos];
    if (vmin == 0.0) {
        if (dec(st.a_min[slot]) == 0.0) atomicMin(st.a_minpos + slot, (int32_t)pos);
        if (dec(st.b_min[o]) == 0.0) atomicMin(st.b_minpos + o, (int32_t)pos);
    }
    if (vmax == 0.0) {
        if (dec(st.a_max[slot]) == 0.0) atomicMin(st.a_maxpos + slot, (int32_t)pos);
        if (dec(st.b_max[o]) == 0.0) atomicMin(st.b_maxpos + o, (int32_t)pos);
    }
}
// ... and the zero of that element, sign included, becomes the extreme
__global__ void __launch_bounds__(BB) k_zero_apply(BuildState st, const int32_t *__restrict__ idx, int64_t n_entries) {
    const int64_t q = (int64_t)blockIdx.x * BB + threadIdx.x;  // entry q < n_active: a slot; then (slot, bucket) pairs
    if (q >= n_entries) return;
    const int64_t n_slots = n_entries / (st.nb + 1);
    const bool is_slot = q < n_slots;
    const int64_t o = is_slot ? q : q - n_slots;
    const int slot = (int)(is_slot ? q : o / st.nb);
    const int dim = st.n_dim[st.active[slot]];
    const int32_t pmin = is_slot ? st.a_minpos[o] : st.b_minpos[o];
    const int32_t pmax = is_slot ? st.a_maxpos[o] : st.b_maxpos[o];
    if (pmin != INT32_MAX) (is_slot ? st.a_min : st.b_min)[o] = enc(st.bb[4 * (int64_t)idx[pmin] + 2 * dim]);
    if (pmax != INT32_MAX) (is_slot ? st.a_max : st.b_max)[o] = enc(st.bb[4 * (int64_t)idx[pmax] + 2 * dim + 1]);
}
__global__ void __launch_bounds__(BB) k_any_negative_zero(const double *__restrict__ v, int64_t n, int32_t *__restrict__ flag) {
    const int64_t i = (int64_t)blockIdx.x * BB + threadIdx.x;
    if (i < n && __double_as_longlong(v[i]) == (long long)0x8000000000000000ULL) *flag = 1;
}

CT_DEV void make_node(const BuildState &st, int id, int ptr, int size, int dim) {  // create_node, creation.py:27-29
    st.n_child[id] = -1;
    st.n_Lmax[id] = -1.0;
    st.n_Rmin[id] = -1.0;
    st.n_ptr[id] = ptr;
    st.n_size[id] = size;
    st.n_dim[id] = (uint8_t)dim;
    st.n_tried[id] = 0;
}

// One thread per active node: everything build() does after the per-bucket bounds (creation.py:308-379).
__global__ void __launch_bounds__(128) k_decide(BuildState st, int n_active) {
    int slot = blockIdx.x * 128 + threadIdx.x;
    if (slot >= n_active) return;
    const int nb = st.nb;
    const int node = st.active[slot];
    const int ptr = st.n_ptr[node], size = st.n_size[node];
    const int dim = st.n_dim[node];
    const int64_t o = (int64_t)slot * nb;
    double range_Rmin = dec(st.a_min[slot]);
    double range_Lmax = dec(st.a_max[slot]);
    double bucket_length = (range_Lmax - range_Rmin) / (double)nb;

    // bucket start offsets for the stable partition
    int run = 0, n_nonempty = 0;
    for (int k = 0; k < nb; k++) {
        st.b_start[o + k] = run;
        int c = st.b_cnt[o + k];
        run += c;
        n_nonempty += (c > 0);
    }
    st.next_left[slot] = -1;
    st.next_right[slot] = -1;
    st.split_pos[slot] = ptr + size;

    if (st.cpl == 1 && size == 2) {  // creation.py:312-320
        st.n_Lmax[node] = range_Lmax;
        st.n_Rmin[node] = range_Rmin;
        int c = atomicAdd(st.counters + 0, 2);
        st.n_child[node] = c;
        make_node(st, c, ptr, 1, !dim);
        make_node(st, c + 1, ptr + 1, 1, !dim);
        return;
    }
    if (n_nonempty <= 1) {  // one bucket holds everything, creation.py:345-358
        if (!st.n_tried[node]) {
            st.n_tried[node] = 1;
            st.n_dim[node] = (uint8_t)(!dim);
            int ns = atomicAdd(st.counters + 1, 1);
            st.next_active[ns] = node;
            st.next_left[slot] = ns;
            st.next_right[slot] = ns;
        }  // else: stays a (possibly oversized) leaf, Lmax = Rmin = -1 already, flipped dim kept
        return;
    }
    // split_plane over the non-empty buckets, creation.py:174-213
    double plane_min_cost = FLOAT_MAX;
    int plane_k = -1;      // original index of the first bucket right of the plane
    int left_count = 0;
    {
        int bbs_in_left = 0;
        int prev = -1;
        for (int k = 0; k < nb; k++) {
            int c = st.b_cnt[o + k];
            if (c == 0) continue;
            if (prev >= 0) {
                bbs_in_left += st.b_cnt[o + prev];
                int bbs_in_right = size - bbs_in_left;
                double cur_Lmax = dec(st.b_max[o + prev]);
                double next_Rmin = dec(st.b_min[o + k]);
                double left_volume = (cur_Lmax - range_Rmin) / bucket_length;
                double right_volume = (range_Lmax - next_Rmin) / bucket_length;
                double plane_cost = left_volume * (double)bbs_in_left + right_volume * (double)bbs_in_right;
                if (plane_cost < plane_min_cost) {
                    plane_min_cost = plane_cost;
                    plane_k = k;
                    left_count = bbs_in_left;
                }
            }
            prev = k;
        }
    }
    if (plane_k < 0) {  // every cost NaN / not below FLOAT_MAX: the reference indexes out of range
        atomicExch(st.counters + 2, CT_ERR_VALUE);
        return;
    }
    double Lmax = FLOAT_MIN, Rmin = FLOAT_MAX;
    for (int k = 0; k < nb; k++) {
        if (st.b_cnt[o + k] == 0) continue;
        if (k < plane_k) {
            double v = dec(st.b_max[o + k]);
            if (v > Lmax) Lmax = v;
        } else {
            double v = dec(st.b_min[o + k]);
            if (v < Rmin) Rmin = v;
        }
    }
    st.n_Lmax[node] = Lmax;
    st.n_Rmin[node] = Rmin;
    int left_size = left_count, right_size = size - left_count;
    int c = atomicAdd(st.counters + 0, 2);
    st.n_child[node] = c;
    make_node(st, c, ptr, left_size, !dim);
    make_node(st, c + 1, ptr + left_size, right_size, !dim);
    st.split_pos[slot] = ptr + left_size;
    if (left_size > st.cpl) {
        int ns = atomicAdd(st.counters + 1, 1);
        st.next_active[ns] = c;
        st.next_left[slot] = ns;
    }
    if (right_size > st.cpl) {
        int ns = atomicAdd(st.counters + 1, 1);
        st.next_active[ns] = c + 1;
        st.next_right[slot] = ns;
    }
}

// one-hot counters of 4 consecutive buckets, scanned over all positions
struct OneHot4 {
    const int32_t *seg;
    const uint16_t *bkt;
    int group;
    __device__ uint4 operator()(int64_t pos) const {
        uint4 r = make_uint4(0, 0, 0, 0);
        if (seg[pos] >= 0) {
            int k = (int)bkt[pos] - 4 * group;
            if (k == 0) r.x = 1;
            else if (k == 1) r.y = 1;
            else if (k == 2) r.z = 1;
            else if (k == 3) r.w = 1;
        }
        return r;
    }
};
struct Add4 {
    __device__ uint4 operator()(const uint4 &a, const uint4 &b) const { return make_uint4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
};

CT_DEV unsigned pick4(const uint4 &v, int k) { return k == 0 ? v.x : (k == 1 ? v.y : (k == 2 ? v.z : v.w)); }

// rank of an element among the same-bucket elements of its node that precede it (stable_partition keeps order)
__global__ void __launch_bounds__(BB) k_rank(BuildState st, const uint4 *__restrict__ scan, int group, int64_t n, int32_t *__restrict__ rank) {
    int64_t pos = (int64_t)blockIdx.x * BB + threadIdx.x;
    if (pos >= n) return;
    int slot = st.seg[pos];
    if (slot < 0) return;
    int k = (int)st.bkt[pos] - 4 * group;
    if (k < 0 || k > 3) return;
    int node = st.active[slot];
    int ptr = st.n_ptr[node];
    rank[pos] = (int32_t)(pick4(scan[pos], k) - pick4(scan[ptr], k));
}

__global__ void __launch_bounds__(BB) k_scatter(BuildState st, const int32_t *__restrict__ idx_in, const int32_t *__restrict__ rank,
                                               int64_t n, int32_t *__restrict__ idx_out, int32_t *__restrict__ seg_out) {
    int64_t pos = (int64_t)blockIdx.x * BB + threadIdx.x;
    if (pos >= n) return;
    int slot = st.seg[pos];
    if (slot < 0) {
        idx_out[pos] = idx_in[pos];
        seg_out[pos] = -1;
        return;
    }
    int node = st.active[slot];
    int k = st.bkt[pos];
    int np = st.n_ptr[node] + st.b_start[(int64_t)slot * st.nb + k] + rank[pos];
    idx_out[np] = idx_in[pos];
    seg_out[np] = (np < st.split_pos[slot]) ? st.next_left[slot] : st.next_right[slot];
}

// ---- at most four buckets: ranks without the scan over all positions -------------------------------------------------
// The rank of an element among the same-bucket elements of its node that precede it is S(pos) - S(ptr of its node), where
// S(pos)[k] counts the bucket-k elements at positions < pos.  S(pos) = (exclusive scan over the per-block counts that
// k_bucket left)[block of pos] + (bucket-k elements before pos within its block): the first is a scan over n / 256 values
// instead of n, the second four ballots per warp.  k_bucket notes the in-block part of S at every node's first position,
// the scatter pass recomputes S for its own position and subtracts.
__global__ void __launch_bounds__(BB) k_scatter4(BuildState st, const int32_t *__restrict__ idx_in, const uint4 *__restrict__ block_prefix,
                                                const uint4 *__restrict__ node_base, int64_t n, int32_t *__restrict__ idx_out,
                                                int32_t *__restrict__ seg_out) {
    __shared__ uint4 s_warp[BB / 32];
    const int64_t pos = (int64_t)blockIdx.x * BB + threadIdx.x;
    const int slot = pos < n ? st.seg[pos] : -1;
    const int k = slot >= 0 ? (int)st.bkt[pos] : -1;
    const uint4 in_block = block_exclusive4(k, &s_warp);
    if (pos >= n) return;
    if (slot < 0) {
        idx_out[pos] = idx_in[pos];
        seg_out[pos] = -1;
        return;
    }
    const int node = st.active[slot];
    const int ptr = st.n_ptr[node];
    const uint4 b = block_prefix[blockIdx.x], first_block = block_prefix[ptr / BB], first_before = node_base[slot];
    const unsigned rank = (pick4(b, k) + pick4(in_block, k)) - (pick4(first_block, k) + pick4(first_before, k));
    const int np = ptr + st.b_start[(int64_t)slot * st.nb + k] + (int)rank;
    idx_out[np] = idx_in[pos];
    seg_out[np] = (np < st.split_pos[slot]) ? st.next_left[slot] : st.next_right[slot];
}

__global__ void __launch_bounds__(BB) k_fill_i32(int32_t *p, int64_t n, int32_t v) {
    int64_t i = (int64_t)blockIdx.x * BB + threadIdx.x;
    if (i < n) p[i] = v;
}
__global__ void __launch_bounds__(BB) k_iota(int32_t *p, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * BB + threadIdx.x;
    if (i < n) p[i] = (int32_t)i;
}

// ---- renumbering to the reference's node order --------------------------------------------------------------
// splits[b] = number of splitting nodes in the subtree of b (children are always created in a later wave)
__global__ void __launch_bounds__(BB) k_subtree_splits(const int32_t *__restrict__ n_child, int lo, int hi, int32_t *__restrict__ splits) {
    int b = lo + blockIdx.x * BB + threadIdx.x;
    if (b >= hi) return;
    int c = n_child[b];
    splits[b] = (c < 0) ? 0 : 1 + splits[c] + splits[c + 1];
}
// rank[b] = pre-order (left first) index of b among the splitting nodes; children of rank r get 1 + 2r, 2 + 2r
__global__ void __launch_bounds__(BB) k_preorder(const int32_t *__restrict__ n_child, const int32_t *__restrict__ splits, int lo, int hi,
                                                 int32_t *__restrict__ rank, int32_t *__restrict__ final_index,
                                                 int32_t *__restrict__ level, int32_t *__restrict__ max_level) {
    int b = lo + blockIdx.x * BB + threadIdx.x;
    if (b >= hi) return;
    int c = n_child[b];
    if (c < 0) return;
    int r = rank[b];
    final_index[c] = 1 + 2 * r;
    final_index[c + 1] = 2 + 2 * r;
    rank[c] = r + 1;
    rank[c + 1] = r + 1 + splits[c];
    int l = level[b] + 1;
    level[c] = l;
    level[c + 1] = l;
    atomicMax(max_level, l);
}
__global__ void __launch_bounds__(BB) k_emit_nodes(BuildState st, const int32_t *__restrict__ rank, const int32_t *__restrict__ final_index,
                                                  int n_nodes, Node32 *__restrict__ out) {
    int b = blockIdx.x * BB + threadIdx.x;
    if (b >= n_nodes) return;
    Node32 nd;
    nd.Lmax = st.n_Lmax[b];
    nd.Rmin = st.n_Rmin[b];
    nd.child = st.n_child[b] < 0 ? -1 : 1 + 2 * rank[b];
    nd.ptr = st.n_ptr[b];
    nd.size = st.n_size[b];
    nd.dim = st.n_dim[b];
    out[final_index[b]] = nd;
}

// ---- query-side acceleration data derived from the finished tree -----------------------------------------
// elem_xy: vertex coordinates per element row (see geometry.cuh: load_polygon)
__global__ void __launch_bounds__(BB) k_elem_coords(const double2 *__restrict__ vertices, const int32_t *__restrict__ elements,
                                                    int64_t count, double2 *__restrict__ xy, int32_t *__restrict__ collision) {
    int64_t k = (int64_t)blockIdx.x * BB + threadIdx.x;
    if (k >= count) return;
    int v = elements[k];
    const double pad = __longlong_as_double((long long)PAD_VERTEX_BITS);
    double2 c = make_double2(pad, pad);
    if (v >= 0) {
        c = vertices[v];
        if ((unsigned long long)__double_as_longlong(c.x) == PAD_VERTEX_BITS) *collision = 1;  // a real vertex looks like padding
    }
    xy[k] = c;
}

// ---- treelets (common.cuh): three binary levels per 128-byte line -------------------------------------------------
// Binary nodes at the seven heap positions of the treelet rooted at `root` (-1: the position does not exist) and
// the left child of each (-1: leaf, or no such position).
__device__ __forceinline__ void treelet_positions(const Node32 *__restrict__ nodes, int64_t n_nodes, int root, int node[7],
                                                  int child[7], int32_t *__restrict__ err) {
#pragma unroll
    for (int p = 1; p < 7; p++) node[p] = -1;
    node[0] = root;
#pragma unroll
    for (int p = 0; p < 7; p++) {
        child[p] = -1;
        if (node[p] < 0) continue;
        int c = nodes[node[p]].child;
        if (c >= 0 && ((int64_t)c + 1 >= n_nodes || c <= node[p])) {  // children always follow their parent
            *err = 1;
            c = -1;
        }
        child[p] = c;
        if (p < 3 && c >= 0) {
            node[2 * p + 1] = c;
            node[2 * p + 2] = c + 1;
        }
    }
}
// number of treelets directly below treelet lo + i
__global__ void __launch_bounds__(BB) k_treelet_count(const Node32 *__restrict__ nodes, int64_t n_nodes, const int32_t *__restrict__ roots,
                                                     int64_t lo, int64_t n, int32_t *__restrict__ cnt, int32_t *__restrict__ err) {
    int64_t i = (int64_t)blockIdx.x * BB + threadIdx.x;
    if (i >= n) return;
    int node[7], child[7];
    treelet_positions(nodes, n_nodes, roots[lo + i], node, child, err);
    int c = 0;
#pragma unroll
    for (int p = 3; p < 7; p++) c += child[p] >= 0 ? 2 : 0;
    cnt[i] = c;
}
// roots of the next treelet level, in slot order behind those of the treelets before this one
__global__ void __launch_bounds__(BB) k_treelet_children(const Node32 *__restrict__ nodes, int64_t n_nodes, int32_t *__restrict__ roots,
                                                        int64_t lo, int64_t n, const int32_t *__restrict__ off, int64_t next_lo,
                                                        int32_t *__restrict__ child_base, int32_t *__restrict__ err) {
    int64_t i = (int64_t)blockIdx.x * BB + threadIdx.x;
    if (i >= n) return;
    int node[7], child[7];
    treelet_positions(nodes, n_nodes, roots[lo + i], node, child, err);
    int64_t base = next_lo + off[i];
    child_base[lo + i] = (int32_t)base;
#pragma unroll
    for (int p = 3; p < 7; p++)
        if (child[p] >= 0) {
            roots[base++] = child[p];
            roots[base++] = child[p] + 1;
        }
}
__global__ void __launch_bounds__(BB) k_treelet_emit(const Node32 *__restrict__ nodes, int64_t n_nodes, const int32_t *__restrict__ bb_indices,
                                                    const int32_t *__restrict__ roots, const int32_t *__restrict__ child_base, int64_t n,
                                                    Treelet *__restrict__ out, int32_t *__restrict__ err) {
    int64_t i = (int64_t)blockIdx.x * BB + threadIdx.x;
    if (i >= n) return;
    int node[7], child[7];
    treelet_positions(nodes, n_nodes, roots[i], node, child, err);
    Treelet t;
    uint32_t meta = 0, child_off = 0;
    int next_child = 0;
#pragma unroll
    for (int p = 0; p < 7; p++) {
        const int q = p + 1;  // slot
        t.slot[p] = make_double2(0.0, 0.0);
        if (node[p] < 0) continue;
        const Node32 nd = nodes[node[p]];
        if (child[p] >= 0) {
            t.slot[p] = make_double2(nd.Lmax, nd.Rmin);
            if (nd.dim) meta |= 1u << q;
            if (p >= 3) {
                child_off |= (uint32_t)next_child << (4 * (p - 3));
                next_child += 2;
            }
        } else {
            int id0 = nd.size > 0 ? bb_indices[nd.ptr] : -1;
            int id1 = nd.size > 1 ? bb_indices[nd.ptr + 1] : -1;
            t.slot[p].x = __longlong_as_double(((long long)(unsigned)nd.size << 32) | (unsigned)nd.ptr);
            t.slot[p].y = __longlong_as_double(((long long)(unsigned)id1 << 32) | (unsigned)id0);
            meta |= 1u << (8 + q);
        }
    }
    t.child_base = child_base[i];
    t.meta = meta;
    t.child_off = child_off;
    t.root_node = roots[i];
    out[i] = t;
}

static int build_treelets(ct_tree *tree, cudaStream_t s) {
    const int64_t n_nodes = tree->n_nodes;
    Scratch<int32_t> roots, child_base, cnt, off, err;
    Scratch<char> tmp;
    CT_CHECK(roots.alloc(n_nodes + 8, s));
    CT_CHECK(child_base.alloc(n_nodes + 8, s));
    CT_CHECK(cnt.alloc(n_nodes + 8, s));
    CT_CHECK(off.alloc(n_nodes + 8, s));
    CT_CHECK(err.alloc(1, s));
    CT_CUDA(cudaMemsetAsync(err.p, 0, 4, s));
    CT_CUDA(cudaMemsetAsync(roots.p, 0, 4, s));  // the first treelet hangs off binary node 0
    size_t tmp_bytes = 0;
    CT_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, cnt.p, off.p, n_nodes + 1, s));
    CT_CHECK(tmp.alloc(tmp_bytes, s));
    int64_t lo = 0, n = 1;
    while (n > 0) {
        k_treelet_count<<<grid_for(n, BB), BB, 0, s>>>(tree->nodes, n_nodes, roots.p, lo, n, cnt.p, err.p);
        CT_LAUNCH_CHECK();
        size_t bytes = tmp_bytes;
        CT_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, bytes, cnt.p, off.p, n, s));
        count_launch(1);
        int32_t last[2] = {0, 0};
        CT_CUDA(cudaMemcpyAsync(&last[0], off.p + (n - 1), 4, cudaMemcpyDeviceToHost, s));
        CT_CUDA(cudaMemcpyAsync(&last[1], cnt.p + (n - 1), 4, cudaMemcpyDeviceToHost, s));
        CT_CUDA(cudaStreamSynchronize(s));
        const int64_t n_next = (int64_t)last[0] + last[1];
        if (lo + n + n_next > n_nodes) {  // every binary node heads at most one treelet
            set_error("malformed tree: the child links do not form a tree");
            return CT_ERR_VALUE;
        }
        k_treelet_children<<<grid_for(n, BB), BB, 0, s>>>(tree->nodes, n_nodes, roots.p, lo, n, off.p, lo + n, child_base.p, err.p);
        CT_LAUNCH_CHECK();
        lo += n;
        n = n_next;
    }
    tree->n_treelets = lo;
    CT_CHECK(dalloc(&tree->treelets, (size_t)lo, s));
    k_treelet_emit<<<grid_for(lo, BB), BB, 0, s>>>(tree->nodes, n_nodes, tree->bb_indices, roots.p, child_base.p, lo, tree->treelets, err.p);
    CT_LAUNCH_CHECK();
    int32_t h_err = 0;
    CT_CUDA(cudaMemcpyAsync(&h_err, err.p, 4, cudaMemcpyDeviceToHost, s));
    CT_CUDA(cudaStreamSynchronize(s));
    if (h_err) {
        set_error("malformed tree: a child index is out of range or does not follow its parent");
        return CT_ERR_VALUE;
    }
    return CT_OK;
}

// ---- entry grid (common.cuh: EntryGrid) ------------------------------------------------------------------------------
// lo[c]: the smallest double whose grid coordinate reaches column c (bisection over the ordered image; grid_coord is
// monotone).  lo[0] = -inf; +inf where no double reaches the column.
__global__ void __launch_bounds__(BB) k_entry_bounds(double vmin, double scale, int bits, double *__restrict__ lo) {
    const int c = blockIdx.x * BB + threadIdx.x;
    if (c >= (1 << bits)) return;
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    if (c == 0) {
        lo[0] = -inf;
        return;
    }
    const uint32_t target = (uint32_t)c << (16 - bits);
    unsigned long long a = enc(-inf), b = enc(inf);  // coord(a) < target (it is 0); the answer lies in (a, b]
    if (grid_coord(inf, vmin, scale) < target) {
        lo[c] = inf;
        return;
    }
    while (b - a > 1) {
        const unsigned long long m = a + (b - a) / 2;
        if (grid_coord(dec(m), vmin, scale) >= target) b = m;
        else a = m;
    }
    lo[c] = dec(b);
}
// one thread per cell: descend while the whole cell (lo < v <= hi in both dimensions) takes the same single child
__global__ void __launch_bounds__(BB) k_entry_table(const Treelet *__restrict__ treelets, const double *__restrict__ lo, int bits,
                                                   uint32_t *__restrict__ handle) {
    const int cell = blockIdx.x * BB + threadIdx.x;
    const int cells = 1 << bits;
    if (cell >= cells * cells) return;
    const int cx = cell & (cells - 1), cy = cell >> bits;
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    const double lo_x = lo[cx], lo_y = lo[cells + cy];
    const double hi_x = cx + 1 < cells ? dec(enc(lo[cx + 1]) - 1) : inf;
    const double hi_y = cy + 1 < cells ? dec(enc(lo[cells + cy + 1]) - 1) : inf;
    const char *base = reinterpret_cast<const char *>(treelets);
    Cursor c;
    cursor_enter(c, base, ROOT_HANDLE);
    while (!cursor_is_leaf(c)) {
        const bool dim = cursor_dim(c);
        const double cell_lo = dim ? lo_y : lo_x, cell_hi = dim ? hi_y : hi_x;
        const double Lmax = c.plane.x, Rmin = c.plane.y;
        // every v <= hi has "v <= Lmax" and not "v >= Rmin"; every v > lo has "v >= Rmin" and not "v <= Lmax"
        const bool left_only = (cell_hi <= Lmax) && (cell_hi < Rmin);
        const bool right_only = (cell_lo >= Rmin) && (cell_lo >= Lmax);
        if (!(left_only || right_only)) break;
        uint32_t left, right;
        cursor_children(c, left, right);
        cursor_descend(c, base, left_only ? left : right);
    }
    handle[cell] = c.handle;
}

static int build_entry_grid(ct_tree *tree, cudaStream_t s) {
    const double wx = tree->bbox[1] - tree->bbox[0], wy = tree->bbox[3] - tree->bbox[2];
    tree->grid_sx = wx > 0 ? 65536.0 / wx : 0.0;
    tree->grid_sy = wy > 0 ? 65536.0 / wy : 0.0;
    const bool usable = tree->grid_sx > 0.0 && tree->grid_sy > 0.0 && tree->grid_sx < FLOAT_MAX && tree->grid_sy < FLOAT_MAX;
    static int forced = -2;
    if (forced == -2) {
        const char *e = getenv("CELLTREE_ENTRY_BITS");  // experiments: 0 switches the grid off
        forced = e ? atoi(e) : -1;
    }
    // about 4 elements per cell, at most 2048 x 2048 cells (16 MB).  Measured on C2 (16.7 M quads), traversal ms per
    // 100 M points: 512^2 3.78, 1024^2 3.61, 2048^2 3.37, 4096^2 (64 MB) 3.31; on the 2 M-triangle Delaunay tree, whose
    // nodes overlap, 128^2 ... 2048^2 all within 1 %.
    int bits = 0;
    while (bits < 11 && ((int64_t)4 << (2 * (bits + 1))) <= tree->n_elem) bits++;
    if (forced >= 0) bits = forced > 12 ? 12 : forced;
    if (!usable || bits < 2) return CT_OK;
    const int cells = 1 << bits;
    CT_CHECK(dalloc(&tree->entry_lo, (size_t)2 * cells, s));
    CT_CHECK(dalloc(&tree->entry_handle, (size_t)cells * cells, s));
    k_entry_bounds<<<grid_for(cells, BB), BB, 0, s>>>(tree->bbox[0], tree->grid_sx, bits, tree->entry_lo);
    CT_LAUNCH_CHECK();
    k_entry_bounds<<<grid_for(cells, BB), BB, 0, s>>>(tree->bbox[2], tree->grid_sy, bits, tree->entry_lo + cells);
    CT_LAUNCH_CHECK();
    k_entry_table<<<grid_for((int64_t)cells * cells, BB), BB, 0, s>>>(tree->treelets, tree->entry_lo, bits, tree->entry_handle);
    CT_LAUNCH_CHECK();
    tree->entry_bits = bits;
    return CT_OK;
}

static int finish_query_data(ct_tree *tree, cudaStream_t s) {
    const int64_t count = tree->n_elem * tree->M;
    CT_CHECK(dalloc(&tree->elem_xy, (size_t)(count > 0 ? count : 1), s));
    {
        Scratch<int32_t> collision;
        CT_CHECK(collision.alloc(1, s));
        CT_CUDA(cudaMemsetAsync(collision.p, 0, sizeof(int32_t), s));
        k_elem_coords<<<grid_for(count, BB), BB, 0, s>>>(tree->vertices, tree->elements, count, tree->elem_xy, collision.p);
        CT_LAUNCH_CHECK();
        int32_t h = 0;
        CT_CUDA(cudaMemcpyAsync(&h, collision.p, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
        CT_CUDA(cudaStreamSynchronize(s));
        tree->length_from_rows = h != 0;
    }
    CT_CHECK(build_treelets(tree, s));
    CT_CHECK(build_entry_grid(tree, s));
    return CT_OK;
}

static int64_t pessimistic_n_nodes(int64_t n_elements) {  // creation.py:216-230
    int64_t n_nodes = n_elements;
    int64_t nodes = (n_elements + 1) / 2;
    while (nodes > 1) {
        n_nodes += nodes;
        nodes = (nodes + 1) / 2;
    }
    return n_nodes + 1;
}

static double now_ms() {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

static int build_tree(ct_tree *tree, cudaStream_t s) {
    const bool debug = getenv("CELLTREE_DEBUG") != nullptr;
    double t_start = now_ms();
    const int64_t n = tree->n_elem;
    const int nb = tree->n_buckets, cpl = tree->cells_per_leaf;
    const int64_t cap_nodes = pessimistic_n_nodes(n) + 2;
    // an active node holds more than cells_per_leaf elements
    const int64_t cap_active = n / (cpl + 1) + 2;

    Scratch<int32_t> idx_a, idx_b, seg_a, seg_b, rank, n_ptr, n_size, n_child, act_a, act_b, b_cnt, b_start, next_left, next_right,
        split_pos, counters, splits, pre_rank, final_index, level;
    Scratch<uint16_t> bkt;
    Scratch<uint8_t> n_dim, n_tried;
    Scratch<double> n_Lmax, n_Rmin;
    Scratch<unsigned long long> a_min, a_max, b_min, b_max;
    Scratch<uint4> scan;
    Scratch<char> scan_tmp;

    CT_CHECK(idx_a.alloc(n, s));
    CT_CHECK(idx_b.alloc(n, s));
    CT_CHECK(seg_a.alloc(n, s));
    CT_CHECK(seg_b.alloc(n, s));
    CT_CHECK(bkt.alloc(n, s));
    if (nb > 4) {  // the scan over all positions and the ranks it yields: only with more than four buckets (see k_scatter4)
        CT_CHECK(rank.alloc(n, s));
        CT_CHECK(scan.alloc(n, s));
    }
    CT_CHECK(n_ptr.alloc(cap_nodes, s));
    CT_CHECK(n_size.alloc(cap_nodes, s));
    CT_CHECK(n_child.alloc(cap_nodes, s));
    CT_CHECK(n_dim.alloc(cap_nodes, s));
    CT_CHECK(n_tried.alloc(cap_nodes, s));
    CT_CHECK(n_Lmax.alloc(cap_nodes, s));
    CT_CHECK(n_Rmin.alloc(cap_nodes, s));
    CT_CHECK(act_a.alloc(cap_active, s));
    CT_CHECK(act_b.alloc(cap_active, s));
    CT_CHECK(a_min.alloc(cap_active, s));
    CT_CHECK(a_max.alloc(cap_active, s));
    // per (active node, bucket) arrays: sized for the waves of a tree with few buckets, grown by a wave that needs more
    // (n_active * n_buckets entries; with thousands of buckets cap_active * n_buckets would not fit any memory)
    int64_t cap_buckets = cap_active * nb < 4 * n + 1024 ? cap_active * nb : 4 * n + 1024;
    CT_CHECK(b_cnt.alloc(cap_buckets, s));
    CT_CHECK(b_min.alloc(cap_buckets, s));
    CT_CHECK(b_max.alloc(cap_buckets, s));
    CT_CHECK(b_start.alloc(cap_buckets, s));
    // at most four buckets: per-block bucket counts, their scan, and the scan value at every active node's first position
    const bool few_buckets = nb <= 4;
    const int64_t n_blocks = grid_for(n, BB);
    Scratch<uint4> block_counts, block_prefix, node_base;
    Scratch<char> block_scan_tmp;
    size_t block_scan_bytes = 0;
    if (few_buckets) {
        CT_CHECK(block_counts.alloc(n_blocks, s));
        CT_CHECK(block_prefix.alloc(n_blocks, s));
        CT_CHECK(node_base.alloc(cap_active, s));
        CT_CUDA(cub::DeviceScan::ExclusiveScan(nullptr, block_scan_bytes, block_counts.p, block_prefix.p, Add4(), make_uint4(0, 0, 0, 0), n_blocks, s));
        CT_CHECK(block_scan_tmp.alloc(block_scan_bytes, s));
    }
    Scratch<int32_t> a_minpos, a_maxpos, b_minpos, b_maxpos;  // signed-zero mode only
    bool signed_zero = false;
    CT_CHECK(next_left.alloc(cap_active, s));
    CT_CHECK(next_right.alloc(cap_active, s));
    CT_CHECK(split_pos.alloc(cap_active, s));
    CT_CHECK(counters.alloc(4, s));

    size_t scan_bytes = 0;
    if (nb > 4) {
        OneHot4 oh{seg_a.p, bkt.p, 0};
        auto in = cub::TransformInputIterator<uint4, OneHot4, cub::CountingInputIterator<int64_t>>(cub::CountingInputIterator<int64_t>(0), oh);
        CT_CUDA(cub::DeviceScan::ExclusiveScan(nullptr, scan_bytes, in, scan.p, Add4(), make_uint4(0, 0, 0, 0), n, s));
    }
    CT_CHECK(scan_tmp.alloc(scan_bytes, s));

    // root: Node(-1, -1.0, -1.0, ptr=0, size=n, dim=False), creation.py:399-401
    k_iota<<<grid_for(n, BB), BB, 0, s>>>(idx_a.p, n);
    CT_LAUNCH_CHECK();
    k_fill_i32<<<grid_for(n, BB), BB, 0, s>>>(seg_a.p, n, n > cpl ? 0 : -1);
    CT_LAUNCH_CHECK();
    {
        int32_t h_ptr = 0, h_size = (int32_t)n, h_child = -1, h_act = 0;
        uint8_t h_zero = 0;
        double h_m1 = -1.0;
        int32_t h_counters[4] = {1, 0, 0, 0};
        CT_CUDA(cudaMemcpyAsync(n_ptr.p, &h_ptr, 4, cudaMemcpyHostToDevice, s));
        CT_CUDA(cudaMemcpyAsync(n_size.p, &h_size, 4, cudaMemcpyHostToDevice, s));
        CT_CUDA(cudaMemcpyAsync(n_child.p, &h_child, 4, cudaMemcpyHostToDevice, s));
        CT_CUDA(cudaMemcpyAsync(n_dim.p, &h_zero, 1, cudaMemcpyHostToDevice, s));
        CT_CUDA(cudaMemcpyAsync(n_tried.p, &h_zero, 1, cudaMemcpyHostToDevice, s));
        CT_CUDA(cudaMemcpyAsync(n_Lmax.p, &h_m1, 8, cudaMemcpyHostToDevice, s));
        CT_CUDA(cudaMemcpyAsync(n_Rmin.p, &h_m1, 8, cudaMemcpyHostToDevice, s));
        CT_CUDA(cudaMemcpyAsync(act_a.p, &h_act, 4, cudaMemcpyHostToDevice, s));
        CT_CUDA(cudaMemcpyAsync(counters.p, h_counters, 16, cudaMemcpyHostToDevice, s));
        // signed-zero mode (k_zero_first) only when some bounding box holds a -0.0
        k_any_negative_zero<<<grid_for(4 * n, BB), BB, 0, s>>>(tree->bb_coords, 4 * n, counters.p + 3);
        CT_LAUNCH_CHECK();
        int32_t h_flag = 0;
        CT_CUDA(cudaMemcpyAsync(&h_flag, counters.p + 3, 4, cudaMemcpyDeviceToHost, s));
        CT_CUDA(cudaStreamSynchronize(s));  // the host temporaries above go out of scope
        signed_zero = h_flag != 0;
    }
    if (signed_zero) {
        CT_CHECK(a_minpos.alloc(cap_active, s));
        CT_CHECK(a_maxpos.alloc(cap_active, s));
        CT_CHECK(b_minpos.alloc(cap_buckets, s));
        CT_CHECK(b_maxpos.alloc(cap_buckets, s));
    }

    double t_alloc = now_ms();
    BuildState st;
    st.bb = tree->bb_coords;
    st.bkt = bkt.p;
    st.n_ptr = n_ptr.p; st.n_size = n_size.p; st.n_child = n_child.p;
    st.n_dim = n_dim.p; st.n_tried = n_tried.p; st.n_Lmax = n_Lmax.p; st.n_Rmin = n_Rmin.p;
    st.a_min = a_min.p; st.a_max = a_max.p;
    st.b_cnt = b_cnt.p; st.b_min = b_min.p; st.b_max = b_max.p; st.b_start = b_start.p;
    st.next_left = next_left.p; st.next_right = next_right.p; st.split_pos = split_pos.p;
    st.counters = counters.p;
    st.a_minpos = a_minpos.p; st.a_maxpos = a_maxpos.p; st.b_minpos = b_minpos.p; st.b_maxpos = b_maxpos.p;
    st.block_counts = block_counts.p;
    st.node_inblock = node_base.p;
    st.nb = nb; st.cpl = cpl;

    int32_t *idx_cur = idx_a.p, *idx_nxt = idx_b.p, *seg_cur = seg_a.p, *seg_nxt = seg_b.p, *act_cur = act_a.p, *act_nxt = act_b.p;
    std::vector<int32_t> wave_start;  // node-count at the start of every wave
    wave_start.push_back(0);
    int n_active = n > cpl ? 1 : 0;
    int32_t node_count = 1;
    const int n_groups = (nb + 3) / 4;
    while (n_active > 0) {
        wave_start.push_back(node_count);
        if ((int64_t)n_active * nb > cap_buckets) {
            cap_buckets = (int64_t)n_active * nb;
            CT_CHECK(b_cnt.alloc(cap_buckets, s));
            CT_CHECK(b_min.alloc(cap_buckets, s));
            CT_CHECK(b_max.alloc(cap_buckets, s));
            CT_CHECK(b_start.alloc(cap_buckets, s));
            st.b_cnt = b_cnt.p; st.b_min = b_min.p; st.b_max = b_max.p; st.b_start = b_start.p;
            if (signed_zero) {
                CT_CHECK(b_minpos.alloc(cap_buckets, s));
                CT_CHECK(b_maxpos.alloc(cap_buckets, s));
                st.b_minpos = b_minpos.p; st.b_maxpos = b_maxpos.p;
            }
        }
        st.seg = seg_cur;
        st.active = act_cur;
        st.next_active = act_nxt;
        k_init_slots<<<grid_for(n_active, BB), BB, 0, s>>>(st, n_active);
        CT_LAUNCH_CHECK();
        k_range<<<grid_for(n, BB), BB, 0, s>>>(st, idx_cur, n);
        CT_LAUNCH_CHECK();
        k_bucket<<<grid_for(n, BB), BB, 0, s>>>(st, idx_cur, n);
        CT_LAUNCH_CHECK();
        if (signed_zero) {
            k_zero_first<<<grid_for(n, BB), BB, 0, s>>>(st, idx_cur, n);
            CT_LAUNCH_CHECK();
            const int64_t entries = (int64_t)n_active * (nb + 1);
            k_zero_apply<<<grid_for(entries, BB), BB, 0, s>>>(st, idx_cur, entries);
            CT_LAUNCH_CHECK();
        }
        k_decide<<<grid_for(n_active, 128), 128, 0, s>>>(st, n_active);
        CT_LAUNCH_CHECK();
        if (few_buckets) {
            size_t bytes = block_scan_bytes;
            CT_CUDA(cub::DeviceScan::ExclusiveScan(block_scan_tmp.p, bytes, block_counts.p, block_prefix.p, Add4(), make_uint4(0, 0, 0, 0), n_blocks, s));
            count_launch(2);
            k_scatter4<<<grid_for(n, BB), BB, 0, s>>>(st, idx_cur, block_prefix.p, node_base.p, n, idx_nxt, seg_nxt);
            CT_LAUNCH_CHECK();
        } else {
            for (int g = 0; g < n_groups; g++) {
                OneHot4 oh{seg_cur, bkt.p, g};
                auto in = cub::TransformInputIterator<uint4, OneHot4, cub::CountingInputIterator<int64_t>>(cub::CountingInputIterator<int64_t>(0), oh);
                size_t bytes = scan_bytes;
                CT_CUDA(cub::DeviceScan::ExclusiveScan(scan_tmp.p, bytes, in, scan.p, Add4(), make_uint4(0, 0, 0, 0), n, s));
                count_launch(2);
                k_rank<<<grid_for(n, BB), BB, 0, s>>>(st, scan.p, g, n, rank.p);
                CT_LAUNCH_CHECK();
            }
            k_scatter<<<grid_for(n, BB), BB, 0, s>>>(st, idx_cur, rank.p, n, idx_nxt, seg_nxt);
            CT_LAUNCH_CHECK();
        }
        int32_t h[4];
        CT_CUDA(cudaMemcpyAsync(h, counters.p, 16, cudaMemcpyDeviceToHost, s));
        CT_CUDA(cudaStreamSynchronize(s));
        if (h[2] == CT_ERR_UNBUCKETABLE) {
            set_error("tree construction: the centroid of an element falls in no bucket (list index out of range)");
            return CT_ERR_UNBUCKETABLE;
        }
        if (h[2] != 0) {
            set_error("tree construction: no split plane has a finite cost (list index out of range)");
            return CT_ERR_UNBUCKETABLE;
        }
        node_count = h[0];
        n_active = h[1];
        int32_t zero = 0;
        CT_CUDA(cudaMemcpyAsync(counters.p + 1, &zero, 4, cudaMemcpyHostToDevice, s));
        CT_CUDA(cudaStreamSynchronize(s));
        std::swap(idx_cur, idx_nxt);
        std::swap(seg_cur, seg_nxt);
        std::swap(act_cur, act_nxt);
        if ((int64_t)wave_start.size() > 4 * (int64_t)n + 64) {
            set_error("tree construction did not terminate");
            return CT_ERR_VALUE;
        }
    }
    wave_start.push_back(node_count);
    double t_waves = now_ms();

    // renumber: creation order -> the reference's depth-first numbering
    CT_CHECK(splits.alloc(node_count, s));
    CT_CHECK(pre_rank.alloc(node_count, s));
    CT_CHECK(final_index.alloc(node_count, s));
    CT_CHECK(level.alloc(node_count, s));
    Scratch<int32_t> max_level;
    CT_CHECK(max_level.alloc(1, s));
    {
        int32_t zero = 0, one = 1;
        CT_CUDA(cudaMemcpyAsync(pre_rank.p, &zero, 4, cudaMemcpyHostToDevice, s));
        CT_CUDA(cudaMemcpyAsync(final_index.p, &zero, 4, cudaMemcpyHostToDevice, s));
        CT_CUDA(cudaMemcpyAsync(level.p, &one, 4, cudaMemcpyHostToDevice, s));
        CT_CUDA(cudaMemcpyAsync(max_level.p, &one, 4, cudaMemcpyHostToDevice, s));
        CT_CUDA(cudaStreamSynchronize(s));
    }
    const int n_waves = (int)wave_start.size() - 1;
    for (int w = n_waves - 1; w >= 0; w--) {
        int lo = wave_start[w], hi = wave_start[w + 1];
        if (hi <= lo) continue;
        k_subtree_splits<<<grid_for(hi - lo, BB), BB, 0, s>>>(n_child.p, lo, hi, splits.p);
        CT_LAUNCH_CHECK();
    }
    for (int w = 0; w < n_waves; w++) {
        int lo = wave_start[w], hi = wave_start[w + 1];
        if (hi <= lo) continue;
        k_preorder<<<grid_for(hi - lo, BB), BB, 0, s>>>(n_child.p, splits.p, lo, hi, pre_rank.p, final_index.p, level.p, max_level.p);
        CT_LAUNCH_CHECK();
    }
    CT_CHECK(dalloc(&tree->nodes, (size_t)node_count, s));
    k_emit_nodes<<<grid_for(node_count, BB), BB, 0, s>>>(st, pre_rank.p, final_index.p, node_count, tree->nodes);
    CT_LAUNCH_CHECK();
    CT_CHECK(dalloc(&tree->bb_indices, (size_t)(n > 0 ? n : 1), s));
    CT_CUDA(cudaMemcpyAsync(tree->bb_indices, idx_cur, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToDevice, s));
    int32_t h_level = 1;
    CT_CUDA(cudaMemcpyAsync(&h_level, max_level.p, 4, cudaMemcpyDeviceToHost, s));
    CT_CUDA(cudaStreamSynchronize(s));
    tree->n_nodes = node_count;
    tree->depth = h_level;
    if (debug)
        fprintf(stderr, "[celltree] build n=%lld: alloc %.2f ms, %d waves %.2f ms, renumber %.2f ms\n", (long long)n,
                t_alloc - t_start, (int)wave_start.size() - 2, t_waves - t_alloc, now_ms() - t_waves);
    return CT_OK;
}

// bbox_tree + default tolerance from the device bb_coords
static int finish_bounds(ct_tree *tree, cudaStream_t s) {
    Scratch<unsigned long long> red;
    CT_CHECK(red.alloc(5, s));
    unsigned long long init[5] = {enc_host(FLOAT_MAX), enc_host(FLOAT_MIN), enc_host(FLOAT_MAX), enc_host(FLOAT_MIN), enc_host(FLOAT_MIN)};
    CT_CUDA(cudaMemcpyAsync(red.p, init, sizeof(init), cudaMemcpyHostToDevice, s));
    int grid = grid_for(tree->n_elem, BB);
    if (grid > 148 * 8) grid = 148 * 8;
    k_bbox_reduce<<<grid, BB, 0, s>>>(tree->bb_coords, tree->n_elem, red.p);
    CT_LAUNCH_CHECK();
    unsigned long long h[5];
    CT_CUDA(cudaMemcpyAsync(h, red.p, sizeof(h), cudaMemcpyDeviceToHost, s));
    CT_CUDA(cudaStreamSynchronize(s));
    for (int k = 0; k < 4; k++) tree->bbox[k] = dec_host(h[k]);
    double diag = dec_host(h[4]);
    double tol = TOLERANCE_FACTOR * diag;  // default_tolerance, celltree_base.py:51-52
    tree->default_tolerance = tol > MIN_TOLERANCE ? tol : MIN_TOLERANCE;
    return CT_OK;
}

static int upload_mesh(ct_tree *tree, const double *vertices, const int64_t *elements, int32_t mem, cudaStream_t s) {
    const int64_t nv = tree->n_vertex, ne = tree->n_elem;
    const int M = tree->M;
    CT_CHECK(dalloc(&tree->vertices, (size_t)(nv > 0 ? nv : 1), s));
    CT_CHECK(dalloc(&tree->elements, (size_t)(ne * M > 0 ? ne * M : 1), s));
    CT_CHECK(dalloc(&tree->bb_coords, 4 * (size_t)(ne > 0 ? ne : 1), s));
    if (mem == CT_MEM_DEVICE) {
        CT_CUDA(cudaMemcpyAsync(tree->vertices, vertices, sizeof(double2) * (size_t)nv, cudaMemcpyDeviceToDevice, s));
        CT_CHECK(launch_narrow(elements, ne * M, tree->elements, s));
    } else {
        CT_CHECK(upload_from_host(tree->vertices, vertices, sizeof(double2) * (size_t)nv, s));
        Scratch<int64_t> tmp;
        CT_CHECK(tmp.alloc((size_t)ne * M, s));
        CT_CHECK(upload_from_host(tmp.p, elements, sizeof(int64_t) * (size_t)ne * M, s));
        CT_CHECK(launch_narrow(tmp.p, ne * M, tree->elements, s));
        CT_CUDA(cudaStreamSynchronize(s));
    }
    return CT_OK;
}

static int check_mesh_args(const double *vertices, int64_t n_vertex, const int64_t *elements, int64_t n_elem, int32_t n_max_vert,
                           int32_t kind) {
    if (!vertices || !elements || n_vertex <= 0 || n_elem <= 0) {
        set_error("zero-size array to reduction operation minimum which has no identity (empty mesh)");
        return CT_ERR_VALUE;
    }
    if (kind == CT_KIND_FACES && (n_max_vert < 3 || n_max_vert > MAX_N_VERTEX)) {
        set_error("faces must have between 3 and 32 columns");
        return CT_ERR_VALUE;
    }
    if (kind == CT_KIND_EDGES && n_max_vert != 2) {
        set_error("edges must have shape (n_edge, 2)");
        return CT_ERR_VALUE;
    }
    if (kind != CT_KIND_FACES && kind != CT_KIND_EDGES) {
        set_error("unknown tree kind");
        return CT_ERR_VALUE;
    }
    if (n_vertex >= (1LL << 31) || n_elem >= (1LL << 31) - 2) {
        set_error("mesh too large for 32-bit device indices");
        return CT_ERR_VALUE;
    }
    return CT_OK;
}

}  // namespace ct

using namespace ct;

extern "C" int ct_tree_create(const double *vertices, int64_t n_vertex, const int64_t *elements, int64_t n_elem, int32_t n_max_vert,
                              int32_t kind, int32_t n_buckets, int32_t cells_per_leaf, double edge_tolerance, int32_t mem,
                              ct_tree **out) {
    if (!out) {
        set_error("ct_tree_create: null output");
        return CT_ERR_VALUE;
    }
    if (n_buckets < 2) {
        set_error("n_buckets must be >= 2");
        return CT_ERR_VALUE;
    }
    if (n_buckets > 65535) {  // the bucket of an element is kept in 16 bits; the reference has no upper bound (celltree.py:69-72)
        set_error("n_buckets must be <= 65535");
        return CT_ERR_VALUE;
    }
    if (cells_per_leaf < 1) {
        set_error("cells_per_leaf must be >= 1");
        return CT_ERR_VALUE;
    }
    CT_CHECK(check_mesh_args(vertices, n_vertex, elements, n_elem, n_max_vert, kind));
    int device = 0;
    CT_CUDA(cudaGetDevice(&device));
    cudaStream_t s = current_stream();
    ct_tree *tree = new ct_tree();
    tree->device = device;
    tree->n_vertex = n_vertex;
    tree->n_elem = n_elem;
    tree->M = n_max_vert;
    tree->kind = kind;
    tree->n_buckets = n_buckets;
    tree->cells_per_leaf = cells_per_leaf;
    auto body = [&]() -> int {
        CT_CHECK(upload_mesh(tree, vertices, elements, mem, s));
        cudaEvent_t e0, e1;
        CT_CUDA(cudaEventCreate(&e0));
        CT_CUDA(cudaEventCreate(&e1));
        CT_CUDA(cudaEventRecord(e0, s));
        if (kind == CT_KIND_FACES) {
            CT_CHECK(launch_counter_clockwise(tree->vertices, tree->elements, n_elem, n_max_vert, s));
            CT_CHECK(launch_face_bboxes(tree->vertices, tree->elements, n_elem, n_max_vert, tree->bb_coords, s));
        } else {
            k_edge_bboxes<<<grid_for(n_elem, BB), BB, 0, s>>>(tree->vertices, tree->elements, n_elem, edge_tolerance, tree->bb_coords);
            CT_LAUNCH_CHECK();
        }
        CT_CHECK(finish_bounds(tree, s));
        CT_CHECK(build_tree(tree, s));
        CT_CHECK(finish_query_data(tree, s));
        CT_CUDA(cudaEventRecord(e1, s));
        CT_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        CT_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        tree->build_ms = ms;
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        return CT_OK;
    };
    int status = body();
    if (status != CT_OK) {
        ct_tree_destroy(tree);
        return status;
    }
    *out = tree;
    return CT_OK;
}

// nodes (41-byte rows, host or device) -> tree->nodes (already allocated, tree->n_nodes rows) and tree->depth
static int unpack_nodes(ct_tree *tree, const ct_node41 *nodes, int32_t mem, cudaStream_t s) {
    const int64_t n_nodes = tree->n_nodes;
    Scratch<unsigned char> packed;
    const size_t bytes = (size_t)n_nodes * NODE41;
    CT_CHECK(packed.alloc(bytes + 4, s));
    if (mem != CT_MEM_DEVICE) CT_CHECK(upload_from_host(packed.p, nodes, bytes, s));
    else CT_CUDA(cudaMemcpyAsync(packed.p, nodes, bytes, cudaMemcpyDeviceToDevice, s));
    k_unpack_nodes<<<grid_for(n_nodes, BB), BB, 0, s>>>(packed.p, n_nodes, tree->nodes);
    CT_LAUNCH_CHECK();
    Scratch<int32_t> parent, max_depth;
    CT_CHECK(parent.alloc(n_nodes, s));
    CT_CHECK(max_depth.alloc(1, s));
    k_fill_i32<<<grid_for(n_nodes, BB), BB, 0, s>>>(parent.p, n_nodes, -1);
    CT_LAUNCH_CHECK();
    CT_CUDA(cudaMemsetAsync(max_depth.p, 0, 4, s));
    k_parents<<<grid_for(n_nodes, BB), BB, 0, s>>>(tree->nodes, n_nodes, parent.p);
    CT_LAUNCH_CHECK();
    k_depth<<<grid_for(n_nodes, BB), BB, 0, s>>>(parent.p, n_nodes, max_depth.p);
    CT_LAUNCH_CHECK();
    int32_t d = 0;
    CT_CUDA(cudaMemcpyAsync(&d, max_depth.p, 4, cudaMemcpyDeviceToHost, s));
    CT_CUDA(cudaStreamSynchronize(s));
    tree->depth = d;
    return CT_OK;
}

// The reference's queries read tree.nodes on every call (query.py:73), so a caller who edits that array changes the
// answers (tests/test_celltree.py:606-618 does).  Here the queries read treelets derived from the nodes: this entry
// replaces the node array by the caller's edited copy and derives the treelets and the entry grid again.  If the new
// links are refused (a child out of range or not following its parent) the tree keeps its previous state.
extern "C" int ct_tree_update_nodes(ct_tree *tree, const ct_node41 *nodes, int64_t n_nodes, int32_t mem) {
    if (!tree || !nodes || n_nodes != tree->n_nodes) {
        set_error("ct_tree_update_nodes: null argument, or a node count other than the tree's");
        return CT_ERR_VALUE;
    }
    CT_ON_DEVICE(tree->device);
    cudaStream_t s = current_stream();
    Node32 *old_nodes = tree->nodes;
    Treelet *old_treelets = tree->treelets;
    uint32_t *old_handle = tree->entry_handle;
    double *old_lo = tree->entry_lo;
    const int64_t old_n_treelets = tree->n_treelets;
    const int32_t old_depth = tree->depth, old_bits = tree->entry_bits;
    tree->nodes = nullptr, tree->treelets = nullptr, tree->entry_handle = nullptr, tree->entry_lo = nullptr;
    auto body = [&]() -> int {
        CT_CHECK(dalloc(&tree->nodes, (size_t)n_nodes, s));
        CT_CHECK(unpack_nodes(tree, nodes, mem, s));
        CT_CHECK(build_treelets(tree, s));
        CT_CHECK(build_entry_grid(tree, s));
        CT_CUDA(cudaStreamSynchronize(s));
        return CT_OK;
    };
    const int status = body();
    if (status != CT_OK) {
        dfree(tree->nodes, s), dfree(tree->treelets, s), dfree(tree->entry_handle, s), dfree(tree->entry_lo, s);
        tree->nodes = old_nodes, tree->treelets = old_treelets, tree->entry_handle = old_handle, tree->entry_lo = old_lo;
        tree->n_treelets = old_n_treelets, tree->depth = old_depth, tree->entry_bits = old_bits;
        return status;
    }
    dfree(old_nodes, s), dfree(old_treelets, s), dfree(old_handle, s), dfree(old_lo, s);
    return CT_OK;
}

extern "C" int ct_tree_from_arrays(const double *vertices, int64_t n_vertex, const int64_t *elements, int64_t n_elem,
                                   int32_t n_max_vert, int32_t kind, const ct_node41 *nodes, int64_t n_nodes,
                                   const int64_t *bb_indices, const double *bb_coords, int32_t cells_per_leaf, int32_t mem,
                                   ct_tree **out) {
    if (!out || !nodes || !bb_indices || !bb_coords || n_nodes <= 0) {
        set_error("ct_tree_from_arrays: null argument");
        return CT_ERR_VALUE;
    }
    CT_CHECK(check_mesh_args(vertices, n_vertex, elements, n_elem, n_max_vert, kind));
    int device = 0;
    CT_CUDA(cudaGetDevice(&device));
    cudaStream_t s = current_stream();
    ct_tree *tree = new ct_tree();
    tree->device = device;
    tree->n_vertex = n_vertex;
    tree->n_elem = n_elem;
    tree->M = n_max_vert;
    tree->kind = kind;
    tree->n_buckets = 0;
    tree->cells_per_leaf = cells_per_leaf;
    tree->n_nodes = n_nodes;
    auto body = [&]() -> int {
        CT_CHECK(upload_mesh(tree, vertices, elements, mem, s));
        auto copy_in = [&](void *dst, const void *src, size_t bytes) -> int {
            if (mem != CT_MEM_DEVICE) return upload_from_host(dst, src, bytes, s);
            CT_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, s));
            return CT_OK;
        };
        CT_CHECK(copy_in(tree->bb_coords, bb_coords, sizeof(double) * 4 * (size_t)n_elem));
        CT_CHECK(dalloc(&tree->bb_indices, (size_t)n_elem, s));
        CT_CHECK(dalloc(&tree->nodes, (size_t)n_nodes, s));
        {
            Scratch<int64_t> tmp;
            const int64_t *src = bb_indices;
            if (mem != CT_MEM_DEVICE) {
                CT_CHECK(tmp.alloc(n_elem, s));
                CT_CHECK(upload_from_host(tmp.p, bb_indices, sizeof(int64_t) * (size_t)n_elem, s));
                src = tmp.p;
            }
            CT_CHECK(launch_narrow(src, n_elem, tree->bb_indices, s));
            CT_CHECK(unpack_nodes(tree, nodes, mem, s));
        }
        CT_CHECK(finish_bounds(tree, s));
        CT_CHECK(finish_query_data(tree, s));
        CT_CUDA(cudaStreamSynchronize(s));
        return CT_OK;
    };
    int status = body();
    if (status != CT_OK) {
        ct_tree_destroy(tree);
        return status;
    }
    *out = tree;
    return CT_OK;
}

extern "C" int ct_tree_get_info(const ct_tree *tree, ct_tree_info *info) {
    if (!tree || !info) {
        set_error("ct_tree_get_info: null argument");
        return CT_ERR_VALUE;
    }
    info->n_vertex = tree->n_vertex;
    info->n_elem = tree->n_elem;
    info->n_nodes = tree->n_nodes;
    info->n_max_vert = tree->M;
    info->kind = tree->kind;
    info->n_buckets = tree->n_buckets;
    info->cells_per_leaf = tree->cells_per_leaf;
    info->depth = tree->depth;
    info->device = tree->device;
    for (int k = 0; k < 4; k++) info->bbox[k] = tree->bbox[k];
    info->default_tolerance = tree->default_tolerance;
    info->build_ms = tree->build_ms;
    return CT_OK;
}

extern "C" int ct_tree_download(const ct_tree *tree, ct_node41 *nodes, int64_t *bb_indices, double *bb_coords, int64_t *elements,
                                double *bb_distances, int32_t mem) {
    if (!tree) {
        set_error("ct_tree_download: null tree");
        return CT_ERR_VALUE;
    }
    CT_ON_DEVICE(tree->device);
    cudaStream_t s = current_stream();
    cudaMemcpyKind kind = mem == CT_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    const int64_t n = tree->n_elem;
    if (nodes) {
        Scratch<unsigned char> packed;
        size_t bytes = (size_t)tree->n_nodes * NODE41;
        CT_CHECK(packed.alloc(bytes + 4, s));
        k_pack_nodes<<<grid_for(tree->n_nodes, BB), BB, 0, s>>>(tree->nodes, tree->n_nodes, packed.p);
        CT_LAUNCH_CHECK();
        CT_CUDA(cudaMemcpyAsync(nodes, packed.p, bytes, kind, s));
        CT_CUDA(cudaStreamSynchronize(s));
    }
    if (bb_indices) {
        Scratch<int64_t> wide;
        int64_t *dst = bb_indices;
        if (mem != CT_MEM_DEVICE) {
            CT_CHECK(wide.alloc(n, s));
            dst = wide.p;
        }
        CT_CHECK(launch_widen(tree->bb_indices, n, dst, s));
        if (mem != CT_MEM_DEVICE) CT_CUDA(cudaMemcpyAsync(bb_indices, dst, sizeof(int64_t) * (size_t)n, cudaMemcpyDeviceToHost, s));
        CT_CUDA(cudaStreamSynchronize(s));
    }
    if (bb_coords) CT_CUDA(cudaMemcpyAsync(bb_coords, tree->bb_coords, sizeof(double) * 4 * (size_t)n, kind, s));
    if (elements) {
        Scratch<int64_t> wide;
        int64_t *dst = elements;
        int64_t count = n * tree->M;
        if (mem != CT_MEM_DEVICE) {
            CT_CHECK(wide.alloc(count, s));
            dst = wide.p;
        }
        CT_CHECK(launch_widen(tree->elements, count, dst, s));
        if (mem != CT_MEM_DEVICE) CT_CUDA(cudaMemcpyAsync(elements, dst, sizeof(int64_t) * (size_t)count, cudaMemcpyDeviceToHost, s));
        CT_CUDA(cudaStreamSynchronize(s));
    }
    if (bb_distances) {
        Scratch<double> dist;
        double *dst = bb_distances;
        if (mem != CT_MEM_DEVICE) {
            CT_CHECK(dist.alloc(3 * (size_t)n, s));
            dst = dist.p;
        }
        k_bb_distances<<<grid_for(n, BB), BB, 0, s>>>(tree->bb_coords, n, dst);
        CT_LAUNCH_CHECK();
        if (mem != CT_MEM_DEVICE) CT_CUDA(cudaMemcpyAsync(bb_distances, dst, sizeof(double) * 3 * (size_t)n, cudaMemcpyDeviceToHost, s));
        CT_CUDA(cudaStreamSynchronize(s));
    }
    CT_CUDA(cudaStreamSynchronize(s));
    return CT_OK;
}

extern "C" void ct_tree_destroy(ct_tree *tree) {
    if (!tree) return;
    cudaStream_t s = current_stream();  // the arrays come from the stream-ordered pool
    dfree(tree->nodes, s);
    dfree(tree->bb_indices, s);
    dfree(tree->bb_coords, s);
    dfree(tree->elements, s);
    dfree(tree->vertices, s);
    dfree(tree->elem_xy, s);
    dfree(tree->treelets, s);
    dfree(tree->entry_handle, s);
    dfree(tree->entry_lo, s);
    delete tree;
}
