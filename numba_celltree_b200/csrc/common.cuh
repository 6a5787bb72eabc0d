// common.cuh -- device tree layout, error plumbing and host helpers shared by the translation units.
//
// Everything in csrc/ is compiled with -fmad=false: the reference (Numba/LLVM) emits no fused
// multiply-adds, and bucket membership / on-edge decisions / pair lists depend on the last bit.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/celltree_b200.h"

namespace ct {

// ---- device image of the reference's nodes: 32 bytes -----------------------------------------------------------
// Reference NodeDType is 41 bytes unaligned (constants.py:91-106); the device copy is repacked to 32 aligned bytes.
// It is what the build produces and ct_tree_download returns (reference numbering); the queries read the treelets
// derived from it (below).
struct __align__(16) Node32 {
    double Lmax;
    double Rmin;
    int32_t child;  // left child; right = child + 1; -1 for a leaf
    int32_t ptr;    // into bb_indices
    int32_t size;
    int32_t dim;    // 0 = x, 1 = y
};
static_assert(sizeof(Node32) == 32, "Node32 must be one 32-byte sector");

// ---- treelet: three binary levels in one 128-byte line ------------------------------------------------------
// The queries walk the tree through TREELETS: the binary node at a level that is a multiple of three, its two
// children and its four grandchildren share one 128-byte line, so a descent of three levels costs ONE dependent
// memory access instead of three, and the up-to-eight treelets below it are stored contiguously.  The line holds
// eight 16-byte slots: slot 0 is the header, slots 1..7 are the binary nodes in heap order (slot q has the children
// 2q and 2q + 1), so the header and the treelet's root share the first 32-byte sector and siblings share a sector.
// A binary node is addressed by the handle (treelet index << 3 | slot): its slot sits at byte 16 * handle of the
// array.  The visiting order of the binary nodes -- what the results depend on -- is untouched; Node32 (reference
// numbering) stays the image that ct_tree_download returns.
//   slot q     inner node: (Lmax, Rmin);  leaf: the bits of {int32 ptr, size, id0, id1} (first two element ids)
//   child_base index of the first treelet below this one
//   meta       bit q: dim of slot q | bit 8 + q: slot q is a leaf
//   child_off  4 bits per bottom slot q = 4..7 (at bit 4 * (q - 4)): its left child is treelet child_base + that,
//              its right child the one after
struct __align__(16) Treelet {
    int32_t child_base;
    uint32_t meta;
    uint32_t child_off;
    int32_t root_node;  // index of the binary node in slot 1 (diagnostics)
    double2 slot[7];    // slot[q - 1]
};
static_assert(sizeof(Treelet) == 128, "a treelet must be one 128-byte line");

// ---- entry grid of the point queries -----------------------------------------------------------------------------
// A uniform grid over the tree's bounding box whose cell (cx, cy) holds the handle of the deepest node that EVERY point
// of the cell reaches by single-child decisions -- at each ancestor exactly one of "x <= Lmax" / "x >= Rmin" holds for the
// whole cell -- so a point query starts there instead of at the root with the reference's visiting order intact (no
// ancestor could have pushed a sibling).  A cell is the set of doubles v with lo[c] < v <= hi[c] whose grid_coord() is c;
// a point that sits exactly on lo[c] (it may lie on a split plane and then goes both ways) starts at the root, and so
// does everything grid_coord() clamps from outside the box or NaN.
struct EntryGrid {
    const uint32_t *handle;  // (cells, cells), row = cy; nullptr: always start at the root
    const double *lo;        // lo_x[cells] then lo_y[cells]: the smallest double of every column / row
    int32_t bits;            // cells = 1 << bits per side
    double xmin, ymin, sx, sy;
};

// 16-bit grid coordinate of a value over [vmin, vmin + 65536 / scale): monotone in v; below the box and NaN -> 0
__device__ __forceinline__ uint32_t grid_coord(double v, double vmin, double scale) {
    double f = (v - vmin) * scale;
    f = f >= 0.0 ? f : 0.0;
    return f < 65535.0 ? (uint32_t)f : 65535u;
}

// Overflow area of the traversal stacks (traverse.cuh: Stack) for trees deeper than the per-thread part: `slots`
// columns of (depth - STACK_CAP) entries, entry k of column c at slab[k * slots + c]; a thread takes a column the first
// time its stack outgrows the per-thread part.  All null for trees that fit (every tree the builder makes of a sane mesh).
struct DeepStacks {
    uint32_t *slab;
    int32_t *state;  // [0] columns handed out, [1] set when a thread found none left
    int32_t slots;
};

struct TreeView {  // passed by value to kernels
    const Node32 *nodes;
    const int32_t *bb_indices;
    const double *bb_coords;  // (n_elem, 4): xmin, xmax, ymin, ymax
    const int32_t *elements;  // (n_elem, M) vertex ids, -1 filled
    const double2 *vertices;
    int32_t M;
    int32_t n_elem;
    double bbox[4];
    // per-element vertex coordinates, (n_elem, M) double2: the polygon of element e is read with one contiguous
    // access instead of a face row followed by M dependent vertex gathers
    const double2 *elem_xy;
    // padding of elem_xy rows is PAD_VERTEX (a NaN with a payload of its own), so the length of a polygon of up to four
    // vertices is read off its coordinates; true only if a real vertex carries that very bit pattern: then the id row decides
    bool length_from_rows;
    const Treelet *treelets;
    EntryGrid entry;
    DeepStacks deep;
};

constexpr int MAX_N_VERTEX = 32;      // constants.py:128
// coordinate of the padding vertices of elem_xy: a quiet NaN whose payload no computation produces
constexpr unsigned long long PAD_VERTEX_BITS = 0x7ff8c0dec0dec0deULL;
constexpr double MIN_TOLERANCE = 1e-15;   // constants.py:140
constexpr double TOLERANCE_FACTOR = 1e-12;  // constants.py:141
constexpr double FLOAT_MAX = 1.7976931348623157e308;   // constants.py:144
constexpr double FLOAT_MIN = -1.7976931348623157e308;  // constants.py:143

// ---- error plumbing ----------------------------------------------------------------------------------
void set_error(const std::string &msg);
cudaStream_t current_stream();
void count_launch(int n = 1);

#define CT_CUDA(expr)                                                                                  \
    do {                                                                                               \
        cudaError_t _e = (expr);                                                                       \
        if (_e != cudaSuccess) {                                                                       \
            char _buf[512];                                                                            \
            snprintf(_buf, sizeof(_buf), "CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__, \
                     __LINE__, cudaGetErrorString(_e));                                                \
            ct::set_error(_buf);                                                                       \
            return CT_ERR_CUDA;                                                                        \
        }                                                                                              \
    } while (0)

#define CT_CHECK(expr)            \
    do {                          \
        int _s = (expr);          \
        if (_s != CT_OK) return _s; \
    } while (0)

#define CT_LAUNCH_CHECK()                 \
    do {                                  \
        ct::count_launch();               \
        CT_CUDA(cudaGetLastError());      \
    } while (0)

// Entry points work on the tree's device and leave the caller's current device as they found it.
struct DeviceGuard {
    int previous = -1;
    cudaError_t error = cudaSuccess;
    explicit DeviceGuard(int device) {
        error = cudaGetDevice(&previous);
        if (error == cudaSuccess && previous != device) error = cudaSetDevice(device);
        else previous = -1;  // nothing to restore
    }
    ~DeviceGuard() {
        if (previous >= 0) cudaSetDevice(previous);
    }
    DeviceGuard(const DeviceGuard &) = delete;
    DeviceGuard &operator=(const DeviceGuard &) = delete;
};
#define CT_ON_DEVICE(device)          \
    ct::DeviceGuard _device_guard(device); \
    CT_CUDA(_device_guard.error)

// Device memory comes from a caching pool inside the library (lib.cu): blocks are rounded to a few size classes per
// octave, freed blocks are kept and handed out again in stream order, so a query call does no cudaMalloc / cudaFree
// (and none of the driver pool's remapping, which cost 20 ms per call for 2 GB of results) once the sizes have been seen.
int pool_alloc(void **p, size_t bytes, cudaStream_t s);
void pool_free(void *p, cudaStream_t s);
// every block freed on `s` so far is safe for any stream (call after synchronising `s`, e.g. before destroying it)
void pool_stream_synced(cudaStream_t s);

template <typename T>
inline int dalloc(T **p, size_t count, cudaStream_t s) {
    *p = nullptr;
    if (count == 0) count = 1;
    return pool_alloc((void **)p, count * sizeof(T), s);
}
inline void dfree(void *p, cudaStream_t s) {
    if (p) pool_free(p, s);
}

// Host -> device copy of a caller's array.  Pageable memory (a plain ndarray) goes through the driver's staging at
// ~10 GB/s; large arrays are therefore copied by a few host threads into two page-locked staging blocks that the copy
// engine drains at the PCIe rate while the threads fill the other one.  Page-locked sources are copied directly.
// The call returns when the source has been consumed (the device copy itself is ordered on `s`).
int upload_from_host(void *dst_device, const void *src_host, size_t bytes, cudaStream_t s);

// Owns a stream-ordered allocation for the duration of a call.
template <typename T>
struct Scratch {
    T *p = nullptr;
    cudaStream_t s = 0;
    Scratch() = default;
    Scratch(const Scratch &) = delete;
    Scratch &operator=(const Scratch &) = delete;
    ~Scratch() { dfree(p, s); }
    int alloc(size_t count, cudaStream_t stream) {
        dfree(p, s);
        s = stream;
        return dalloc(&p, count, stream);
    }
    T *release() {
        T *r = p;
        p = nullptr;
        return r;
    }
    operator T *() const { return p; }
};

// CELLTREE_DEBUG=1: synchronise and print the host time since the previous trace point (stderr)
void trace_point(cudaStream_t s, const char *label);

inline int grid_for(int64_t n, int block) {
    int64_t g = (n + block - 1) / block;
    if (g < 1) g = 1;
    return (int)g;
}

// ---- the device tree ---------------------------------------------------------------------------------
}  // namespace ct
struct ct_tree;
struct ct_result;
namespace ct {
// ---- shared host helpers (lib.cu / build.cu) --------------------------------------------------------------
constexpr int BLOCK = 128;  // query kernels: one query (or pair) per thread

// Device view of a caller array: for CT_MEM_HOST a stream-ordered scratch copy, else the pointer itself.
template <typename T>
struct DevIn {
    Scratch<T> owned;
    const T *p = nullptr;
    int init(const T *src, size_t count, int mem, cudaStream_t s) {
        if (mem == CT_MEM_DEVICE) {
            p = src;
            return CT_OK;
        }
        CT_CHECK(owned.alloc(count, s));
        if (count) CT_CHECK(upload_from_host(owned.p, src, count * sizeof(T), s));
        p = owned.p;
        return CT_OK;
    }
};

template <typename T>
struct DevOut {
    Scratch<T> owned;
    T *p = nullptr;
    T *host = nullptr;
    size_t count = 0;
    int init(T *dst, size_t n, int mem, cudaStream_t s) {
        count = n;
        if (dst == nullptr) return CT_OK;
        if (mem == CT_MEM_DEVICE) {
            p = dst;
            return CT_OK;
        }
        host = dst;
        CT_CHECK(owned.alloc(n, s));
        p = owned.p;
        return CT_OK;
    }
    int finish(cudaStream_t s) {
        if (host && count) CT_CUDA(cudaMemcpyAsync(host, p, count * sizeof(T), cudaMemcpyDeviceToHost, s));
        return CT_OK;
    }
};


// Host side of DeepStacks: nothing for trees of at most STACK_CAP levels.  `threads` bounds how many threads of one
// launch can need a column at the same time.
struct DeepScope {
    Scratch<uint32_t> slab;
    Scratch<int32_t> state;
    DeepStacks view{nullptr, nullptr, 0};
    cudaStream_t stream = 0;
    int init(const ct_tree *tree, int64_t threads, cudaStream_t s);
    int next_launch();  // before every launch that walks the tree: all columns are free again
    int finish();       // after the last one: CT_ERR_DEPTH if a thread found no column
};
// phase events of the last ct_locate_points call (lib.cu); nullptr when profiling is off
struct PhaseEvents {
    cudaEvent_t start, ordered, done;
};
PhaseEvents *phase_events();
int sort_bits_override();  // -1 = automatic
// exclusive scan of int32 counts[0..n] (counts[n] must be 0) into int64 offsets[0..n]; *total = offsets[n]
int scan_counts(const int32_t *counts, int64_t n, int64_t *offsets, int64_t *total, cudaStream_t s);
// keep the flagged pairs of a result, order preserved
int compact_result(ct_result *r, const int32_t *flag, bool keep_payload, cudaStream_t s);
int launch_widen(const int32_t *in, int64_t n, int64_t *out, cudaStream_t s);
// int64 -> int32; entries equal to `fill` become -1 (cast_faces, cast.py:38-39)
int launch_narrow(const int64_t *in, int64_t n, int32_t *out, cudaStream_t s, int64_t fill = -1);
int launch_counter_clockwise(const double2 *vertices, int32_t *faces, int64_t n_face, int M, cudaStream_t s);
int launch_face_bboxes(const double2 *vertices, const int32_t *faces, int64_t n_face, int M, double *bb, cudaStream_t s);
}  // namespace ct

struct ct_tree {
    int64_t n_vertex = 0, n_elem = 0, n_nodes = 0;
    int32_t M = 0, kind = 0, n_buckets = 0, cells_per_leaf = 0, depth = 0;
    double bbox[4] = {0, 0, 0, 0};
    double default_tolerance = 0.0;
    double build_ms = 0.0;
    int device = 0;
    // device arrays (owned)
    ct::Node32 *nodes = nullptr;
    int32_t *bb_indices = nullptr;
    double *bb_coords = nullptr;
    int32_t *elements = nullptr;
    double2 *vertices = nullptr;
    double2 *elem_xy = nullptr;
    bool length_from_rows = false;
    ct::Treelet *treelets = nullptr;
    int64_t n_treelets = 0;
    uint32_t *entry_handle = nullptr;
    double *entry_lo = nullptr;
    int32_t entry_bits = 0;
    double grid_sx = 0.0, grid_sy = 0.0;  // 65536 / bbox width, height (0 when the box is degenerate)

    ct::TreeView view() const {
        ct::TreeView v;
        v.nodes = nodes;
        v.bb_indices = bb_indices;
        v.bb_coords = bb_coords;
        v.elements = elements;
        v.vertices = vertices;
        v.M = M;
        v.n_elem = (int32_t)n_elem;
        for (int k = 0; k < 4; k++) v.bbox[k] = bbox[k];
        v.elem_xy = elem_xy;
        v.length_from_rows = length_from_rows;
        v.treelets = treelets;
        v.entry.handle = entry_handle;
        v.entry.lo = entry_lo;
        v.entry.bits = entry_bits;
        v.entry.xmin = bbox[0];
        v.entry.ymin = bbox[2];
        v.entry.sx = grid_sx;
        v.entry.sy = grid_sy;
        v.deep.slab = nullptr;
        v.deep.state = nullptr;
        v.deep.slots = 0;
        return v;
    }
};

struct ct_result {
    int64_t size = 0;
    int32_t width = 0;   // payload doubles per pair
    int32_t *i = nullptr;  // query index
    int32_t *j = nullptr;  // tree element index
    double *payload = nullptr;
};
