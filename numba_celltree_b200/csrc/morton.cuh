// morton.cuh -- Z-order execution order for a batch of box or segment queries (and, as CELLTREE_ORDER=morton, of point
// queries: their default order is the slab binning of binning.cuh, which needs no sort).
//
// Queries are independent, so the order in which threads pick them up is an execution detail: thread t handles
// query perm[t] and writes result slot perm[t] (for variable-length results: counts[perm[t]] in the count pass,
// offsets[perm[t]] in the fill pass -- the OUTPUT order is untouched).  Sorting the queries along a Z-order curve
// over the tree's bounding box makes the 32 lanes of a warp walk (almost) the same root-to-leaf paths, so node /
// face / vertex loads collapse to a few sectors per warp and the lower tree levels are served by L1/L2 instead
// of HBM.  The sort is a CUB radix sort of (key, index) pairs over as many 8-bit passes as the batch size needs.
#pragma once

#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace ct {

enum { KEY_POINT = 0, KEY_BOX = 1, KEY_EDGE = 2 };

__device__ __forceinline__ uint32_t spread16(uint32_t v) {  // 16 bits -> every other bit of 32
    v &= 0xffffu;
    v = (v | (v << 8)) & 0x00ff00ffu;
    v = (v | (v << 4)) & 0x0f0f0f0fu;
    v = (v | (v << 2)) & 0x33333333u;
    v = (v | (v << 1)) & 0x55555555u;
    return v;
}

// representative point of query i: the point itself, the centre of a box (xmin, xmax, ymin, ymax), the midpoint
// of a segment ((x0, y0), (x1, y1))
// For boxes and segments the key's top `class_bits` bits are a SIZE CLASS (extent of the query in units of
// `class_unit`, saturating): the work of such a query grows with its extent, a warp takes as long as its slowest lane,
// so queries of similar extent are put in the same warps; below the class the key is the Z-order of the centre.
template <int KIND>
__global__ void __launch_bounds__(256) k_morton_keys(const double *__restrict__ q, int64_t n, double xmin, double ymin, double sx,
                                                     double sy, int shift, int class_bits, double class_unit,
                                                     uint32_t *__restrict__ keys, uint32_t *__restrict__ idx) {
    int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    double x, y, extent = 0.0;
    if (KIND == KEY_POINT) {
        double2 p = __ldg(reinterpret_cast<const double2 *>(q) + i);
        x = p.x;
        y = p.y;
    } else {
        const double2 *r = reinterpret_cast<const double2 *>(q) + 2 * i;
        double2 a = __ldg(r), b = __ldg(r + 1);
        if (KIND == KEY_BOX) {
            x = 0.5 * (a.x + a.y);
            y = 0.5 * (b.x + b.y);
            extent = fmax(a.y - a.x, b.y - b.x);
        } else {
            x = 0.5 * (a.x + b.x);
            y = 0.5 * (a.y + b.y);
            extent = fabs(b.x - a.x) + fabs(b.y - a.y);
        }
    }
    const uint32_t ix = grid_coord(x, xmin, sx), iy = grid_coord(y, ymin, sy);  // [0, 65536) inside the tree's bounding box
    uint32_t key = (spread16(ix) | (spread16(iy) << 1)) >> shift;
    if (KIND != KEY_POINT && class_bits > 0) {
        const double c = extent * class_unit;  // NaN -> class 0
        const uint32_t top = (1u << class_bits) - 1u;
        const uint32_t size_class = c >= 0.0 ? (c < (double)top ? (uint32_t)c : top) : 0u;
        const int key_bits = 32 - shift;
        key = (key >> class_bits) | (size_class << (key_bits - class_bits));
    }
    keys[i] = key;
    idx[i] = (uint32_t)i;
}

// Number of Morton key bits to sort the queries by; 0 = keep the caller's order.
// Sorting pays when the tree does not sit in L1/L2 anyway and there are enough queries to amortise the passes.
inline int sort_bits_for(const ct_tree *tree, int64_t n) {
    const int forced = sort_bits_override();
    if (n >= (1LL << 31)) return 0;
    if (forced >= 0) return forced > 32 ? 32 : forced;
    const double tree_bytes = 32.0 * (double)tree->n_nodes + (double)tree->n_elem * (4.0 + 20.0 * tree->M);
    if (n < (1 << 17) || tree_bytes < 8e6) return 0;
    int bits = 8;  // about one key per query, whole 8-bit radix passes, at most three of them
    while (bits < 24 && (1LL << bits) < n) bits += 8;
    return bits;
}

struct MortonOrder {
    Scratch<uint32_t> keys_a, keys_b, idx_a, idx_b;
    Scratch<char> tmp;
    const uint32_t *perm = nullptr;  // nullptr: run in the caller's order
    // after build(): the buffer that holds perm, and the three n-element buffers the sort no longer needs
    uint32_t *perm_buffer = nullptr, *spare[3] = {nullptr, nullptr, nullptr};

    template <int KIND>
    int build(const ct_tree *tree, const double *q, int64_t n, cudaStream_t s) {
        perm = nullptr;
        const int bits = sort_bits_for(tree, n);
        if (bits <= 0 || n <= 0) return CT_OK;
        CT_CHECK(keys_a.alloc(n, s));
        CT_CHECK(keys_b.alloc(n, s));
        CT_CHECK(idx_a.alloc(n, s));
        CT_CHECK(idx_b.alloc(n, s));
        const double sx = tree->grid_sx, sy = tree->grid_sy;
        // size classes (segments only, four of them, in units of two mean cell sizes): measured on C3 / C4,
        // intersect_edges 67.3 -> 63.6 ms with 2 class bits; boxes lose (12.7 -> 13.2 ms with 2 bits, 16.1 with 3):
        // their walk is short enough for the spatial order to matter more than the balance
        const double area = (tree->bbox[1] - tree->bbox[0]) * (tree->bbox[3] - tree->bbox[2]);
        const double cell = (area > 0.0 && tree->n_elem > 0) ? sqrt(area / (double)tree->n_elem) : 0.0;
        const int class_bits = (KIND == KEY_EDGE && cell > 0.0 && bits >= 16) ? 2 : 0;
        const double unit = cell > 0.0 ? 0.5 / cell : 0.0;
        k_morton_keys<KIND><<<grid_for(n, 256), 256, 0, s>>>(q, n, tree->bbox[0], tree->bbox[2], sx, sy, 32 - bits, class_bits, unit,
                                                             keys_a.p, idx_a.p);
        CT_LAUNCH_CHECK();
        cub::DoubleBuffer<uint32_t> d_keys(keys_a.p, keys_b.p);
        cub::DoubleBuffer<uint32_t> d_vals(idx_a.p, idx_b.p);
        size_t bytes = 0;
        CT_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, d_keys, d_vals, n, 0, bits, s));
        CT_CHECK(tmp.alloc(bytes, s));
        CT_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, bytes, d_keys, d_vals, n, 0, bits, s));
        count_launch(1 + (bits + 7) / 8);
        perm = d_vals.Current();
        perm_buffer = d_vals.Current();
        spare[0] = d_keys.Current();
        spare[1] = d_keys.Alternate();
        spare[2] = d_vals.Alternate();
        return CT_OK;
    }

    // out[perm[t]] = results[t] for results computed in execution order (results = spare[0], int32).
    // Written directly, these are 8-byte stores all over `out`: every one dirties a 32-byte sector that DRAM must
    // read, merge and write back.  Instead the (index, result) pairs first go through ONE radix pass on the top
    // eight bits of the index, which groups them into 256 contiguous windows of `out`; the stores of a window then
    // meet in L2 and leave it as whole sectors.  perm is consumed.
    int scatter_results(int64_t n, int64_t *out, cudaStream_t s);
};

static __global__ void __launch_bounds__(256) k_scatter_results(const uint32_t *__restrict__ index, const uint32_t *__restrict__ value, int64_t n,
                                                         int64_t *__restrict__ out) {
    int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (t >= n) return;
    out[__ldcs(index + t)] = (int64_t)(int32_t)__ldcs(value + t);
}

inline int MortonOrder::scatter_results(int64_t n, int64_t *out, cudaStream_t s) {
    int top = 1;
    while (top < 32 && ((int64_t)1 << top) < n) top++;
    const int low = top > 8 ? top - 8 : 0;
    cub::DoubleBuffer<uint32_t> d_keys(perm_buffer, spare[2]);
    cub::DoubleBuffer<uint32_t> d_vals(spare[0], spare[1]);
    size_t bytes = 0;
    CT_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, d_keys, d_vals, n, low, top, s));
    Scratch<char> tmp2;
    CT_CHECK(tmp2.alloc(bytes, s));
    CT_CUDA(cub::DeviceRadixSort::SortPairs(tmp2.p, bytes, d_keys, d_vals, n, low, top, s));
    count_launch(2);
    k_scatter_results<<<grid_for(n, 256), 256, 0, s>>>(d_keys.Current(), d_vals.Current(), n, out);
    CT_LAUNCH_CHECK();
    perm = nullptr;
    return CT_OK;
}

}  // namespace ct
