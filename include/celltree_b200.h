/*
 * celltree_b200.h -- C-ABI of libcelltree_b200.so, the sm_100a implementation of the
 * numba_celltree query hot path.
 *
 * The reference (Deltares/numba_celltree, pure Python + Numba) has no FFI layer; its seam is
 * "API class method -> module-level @njit array function" (SURVEY.md section 8b).  Every entry
 * point below replaces one or more of those array functions; the replaced reference interface is
 * cited as file:line (paths relative to the reference's numba_celltree/ package).  The host-side
 * Python classes in numba_celltree_b200/ bind these with ctypes; INTEGRATION.md shows the stub a
 * reference maintainer would add.
 *
 * Conventions
 *   - plain pointers and sizes only; all functions return 0 on success, a CT_ERR_* code otherwise
 *     (ct_last_error() gives the message for the calling thread).
 *   - `mem` says where EVERY array argument of that call lives: CT_MEM_HOST (NumPy buffers; the
 *     call does the host<->device copies) or CT_MEM_DEVICE (pointers into this process' CUDA
 *     context, e.g. torch tensors; no PCIe traffic).  Outputs are written to caller-owned buffers.
 *   - calls are synchronous with respect to the host for CT_MEM_HOST.  For CT_MEM_DEVICE the work
 *     is enqueued on the stream set with ct_set_stream() (default: the legacy default stream) and
 *     only entry points that must return a size (variable-length results) synchronise.
 *   - integer arrays are int64 (np.intp, constants.py:27), floats are float64, C-contiguous.
 *   - box rows are (xmin, xmax, ymin, ymax); edge rows are ((x0, y0), (x1, y1)).
 */
#ifndef CELLTREE_B200_H
#define CELLTREE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CT_OK 0
#define CT_ERR_CUDA 1          /* CUDA runtime error (no device, out of memory, launch failure) */
#define CT_ERR_VALUE 2         /* invalid argument (maps to Python ValueError) */
#define CT_ERR_UNBUCKETABLE 3  /* tree build: an element's centroid falls in no bucket (reference: IndexError, creation.py:130-131) */
#define CT_ERR_DEPTH 4         /* more queries of one call overflowed the per-thread traversal stack than the overflow slab holds */
#define CT_ERR_CLIP_STATE 5    /* "Undefined clipping state" (cohen_sutherland.py:86) */
#define CT_ERR_ZERO_DIVISION 6 /* barycentric weights: a division by exactly zero, where the reference (Numba, Python error
                                  model) raises ZeroDivisionError: barycentric_wachspress.py:34,76,83, barycentric_triangle.py:36 */

#define CT_MEM_HOST 0
#define CT_MEM_DEVICE 1

#define CT_KIND_FACES 0 /* CellTree2d: elements are polygons of <= 32 vertices, -1 filled */
#define CT_KIND_EDGES 1 /* EdgeCellTree2d: elements are 2-vertex segments */

typedef struct ct_tree ct_tree;     /* device-resident cell tree (opaque) */
typedef struct ct_result ct_result; /* device-resident variable-length result (opaque) */

/* Byte-exact image of one row of the reference's NodeDType (constants.py:91-106), 41 bytes. */
#pragma pack(push, 1)
typedef struct ct_node41 {
    int64_t child; /* index of left child; right child is child + 1; -1 for a leaf */
    double Lmax;
    double Rmin;
    int64_t ptr;   /* into bb_indices */
    int64_t size;
    uint8_t dim;   /* 0 = x, 1 = y */
} ct_node41;
#pragma pack(pop)

typedef struct ct_tree_info {
    int64_t n_vertex;
    int64_t n_elem;
    int64_t n_nodes;
    int32_t n_max_vert;
    int32_t kind;
    int32_t n_buckets;
    int32_t cells_per_leaf;
    int32_t depth;            /* number of node levels (root alone = 1) */
    int32_t device;           /* CUDA device the tree lives on */
    double bbox[4];           /* bbox_tree(), celltree_base.py:20-25 */
    double default_tolerance; /* default_tolerance(bb_distances[:, 2]), celltree_base.py:51-52 */
    double build_ms;          /* device time of the build (CUDA events) */
} ct_tree_info;

/* ---- library ------------------------------------------------------------------------------- */
const char *ct_last_error(void);
int ct_device_count(int *count);
int ct_set_device(int device);
/* cudaStream_t as void*; applies to subsequent calls of the calling thread. NULL = legacy default stream. */
int ct_set_stream(void *cuda_stream);
/* number of kernel launches issued by this library since load (bench.py's gpu_launches) */
int64_t ct_launch_count(void);
/* Morton ordering of point queries (an execution detail, results are unaffected): number of Z-order key bits
 * the queries are radix-sorted by before the traversal; 0 = keep the caller's order, -1 = automatic (default;
 * also settable through the environment variable CELLTREE_SORT_BITS). */
int ct_set_sort_bits(int32_t bits);
/* Phase timing of ct_locate_points with CT_MEM_DEVICE (CUDA events on the launching stream): when enabled, every
 * call records events around the Morton ordering and around the traversal kernel; ct_profile_last() waits for the
 * last call and returns both durations in milliseconds.  bench.py's roofline line uses the traversal figure. */
int ct_profile_enable(int32_t enable);
int ct_profile_last(double *order_ms, double *traverse_ms);

/* Page-locked host memory for results (the reference returns freshly allocated ndarrays, query.py:46-59; a
 * device-to-host copy into fresh pageable memory runs at a fraction of the PCIe rate).  Blocks are recycled
 * through a cache inside the library (limit: CELLTREE_PINNED_CACHE_MB, default 8192); ct_host_trim() empties it. */
/* Device memory (trees, scratch, results) is cached inside the library in size classes, so that query calls do not
 * allocate once their sizes have been seen (limit on cached free bytes: CELLTREE_DEVICE_CACHE_MB, default 65536).
 * ct_device_trim() returns every cached free block to the driver. */
void ct_device_trim(void);
int ct_host_alloc(size_t bytes, void **out);
void ct_host_free(void *block);
void ct_host_trim(void);

/* ---- construction --------------------------------------------------------------------------
 * ct_tree_create replaces the constructor pipeline of CellTree2d.__init__ (celltree.py:74-97) /
 * EdgeCellTree2d.__init__ (edge_celltree.py:57-83):
 *   counter_clockwise      geometry_utils.py:541-561   (faces only; the corrected faces are kept)
 *   build_face_bboxes      geometry_utils.py:443-456   / build_edge_bboxes :473-487
 *   creation.initialize    creation.py:384-413         (nodes and bb_indices bit-exact)
 *   bbox_tree, bbox_distances, default_tolerance       celltree_base.py:20-52
 * elements: (n_elem, n_max_vert) int64, FILL_VALUE (-1) padded (cast_faces already applied).
 * edge_tolerance: the bounding-box padding for CT_KIND_EDGES (edge_celltree.py:59-64); ignored for faces.
 */
int ct_tree_create(const double *vertices, int64_t n_vertex, const int64_t *elements, int64_t n_elem,
                   int32_t n_max_vert, int32_t kind, int32_t n_buckets, int32_t cells_per_leaf,
                   double edge_tolerance, int32_t mem, ct_tree **out);

/* Upload a tree that was built elsewhere (arrays as the reference's CellTreeData holds them,
 * constants.py:82-89): no build kernels run. `elements` are used as given (no counter_clockwise). */
int ct_tree_from_arrays(const double *vertices, int64_t n_vertex, const int64_t *elements, int64_t n_elem,
                        int32_t n_max_vert, int32_t kind, const ct_node41 *nodes, int64_t n_nodes,
                        const int64_t *bb_indices, const double *bb_coords, int32_t cells_per_leaf,
                        int32_t mem, ct_tree **out);

int ct_tree_get_info(const ct_tree *tree, ct_tree_info *info);

/* Host/device mirrors of the tree attributes (celltree.py:80-97). Any pointer may be NULL (skipped).
 * nodes: n_nodes x 41 bytes; bb_indices: n_elem; bb_coords: n_elem x 4; elements: n_elem x n_max_vert
 * (after counter_clockwise); bb_distances: n_elem x 3 (dx, dy, diagonal). */
int ct_tree_download(const ct_tree *tree, ct_node41 *nodes, int64_t *bb_indices, double *bb_coords,
                     int64_t *elements, double *bb_distances, int32_t mem);

/* Replace the tree's node array by an edited copy (n_nodes must be the tree's own) and derive the traversal structures
 * again.  The reference's queries read tree.nodes on every call (query.py:73), so editing that array changes the answers
 * (tests/test_celltree.py:606-618); the Python classes call this entry when the caller has modified the `nodes` mirror.
 * Links that are out of range or do not follow their parent are refused (CT_ERR_VALUE) and the tree is left unchanged. */
int ct_tree_update_nodes(ct_tree *tree, const ct_node41 *nodes, int64_t n_nodes, int32_t mem);

void ct_tree_destroy(ct_tree *tree);

/* ---- fixed-size queries ---------------------------------------------------------------------
 * ct_locate_points replaces query.locate_points (query.py:110-117) for CT_KIND_FACES and
 * query.locate_points_on_edge (query.py:168-174) for CT_KIND_EDGES.
 * If weights != NULL (faces only) it also replaces barycentric_triangle_weights
 * (algorithms/barycentric_triangle.py:46-64, n_max_vert == 3) / barycentric_wachspress_weights
 * (algorithms/barycentric_wachspress.py:88-107, n_max_vert > 3): weights is (n, n_max_vert).
 * points: (n, 2); out_index: n (-1 = not found).
 */
int ct_locate_points(const ct_tree *tree, const double *points, int64_t n, double tolerance,
                     int64_t *out_index, double *weights, int32_t mem);

/* ---- variable-length queries: count -> scan -> fill on the device, then fetch ---------------
 * ct_locate_boxes replaces query.locate_boxes (query.py:278-289); with with_area != 0 it also runs
 * algorithms.box_area_of_intersection (sutherland_hodgman.py:171-187) and keeps area > 0
 * (celltree.py:174-184): the payload is then one double per pair (the area).
 */
int ct_locate_boxes(const ct_tree *tree, const double *boxes, int64_t n, int32_t with_area, int32_t mem,
                    ct_result **out);

/* ct_locate_faces replaces CellTree2d.locate_faces' kernels (celltree.py:212-226): counter_clockwise on
 * the QUERY faces, build_face_bboxes, locate_boxes, polygons_intersect (separating_axis.py:58-75).  With
 * with_area != 0 it continues with area_of_intersection (sutherland_hodgman.py:151-168) and keeps area > 0
 * (celltree.py:258-269).
 * fill_value: entries of `faces` equal to it count as padding (cast_faces' rewrite to -1, cast.py:38-39, done
 * on the device so that the caller's array need not be copied on the host first).
 * write_back != 0: `faces` is rewritten with the counter-clockwise faces, as locate_faces does to its argument
 * (celltree.py:212); 0 leaves it untouched (intersect_faces works on a copy, celltree.py:256). */
int ct_locate_faces(const ct_tree *tree, const double *vertices, int64_t n_vertex, int64_t *faces,
                    int64_t n_face, int32_t n_max_vert, int64_t fill_value, int32_t write_back,
                    int32_t with_area, int32_t mem, ct_result **out);

/* ct_intersect_edges replaces query.locate_edge_faces (query.py:542-544, CT_KIND_FACES) /
 * query.locate_edge_edges (query.py:537-539, CT_KIND_EDGES) followed by
 * geometry_utils.sort_intersections_by_edge (geometry_utils.py:564-574).
 * Payload per pair: 4 doubles ((cx, cy), (dx, dy)). */
int ct_intersect_edges(const ct_tree *tree, const double *edges, int64_t n, int32_t mem, ct_result **out);
/* Slots per query segment in the hit log of ct_intersect_edges (default and maximum 32; also CELLTREE_HIT_LOG; negative:
 * back to the default).  The clip of every candidate cell is computed once: hits are counted and kept in their segment's
 * slots, then moved to their place after the scan of the counts; a segment with more hits than slots takes a second
 * traversal instead (0 = every segment).  An execution detail: results are unaffected. */
int ct_set_hit_log(int64_t hits_per_query);

int64_t ct_result_size(const ct_result *result);
int32_t ct_result_payload_width(const ct_result *result); /* doubles per pair: 0, 1 or 4 */
/* Copy out the pairs: i = query index, j = tree element index (both int64), payload may be NULL. */
int ct_result_fetch(const ct_result *result, int64_t *i, int64_t *j, double *payload, int32_t mem);
void ct_result_free(ct_result *result);

/* ---- geometry helpers the reference exports beside the query API ----------------------
 * The reference's versions are scalar @njit functions on Point / Box tuples; these take n inputs per call.
 * a, b, c, d: (n, 2) doubles; intersects / inside: n bytes (0 / 1); c, d are NaN where intersects is 0.
 * boxes: (n_boxes, 4) rows (xmin, xmax, ymin, ymax) with n_boxes == n (one box per segment) or 1 (one box for all).
 *   ct_liang_barsky_line_box_clip      algorithms/liang_barsky.py:10-64
 *   ct_cohen_sutherland_line_box_clip  algorithms/cohen_sutherland.py:36-101
 *   ct_cyrus_beck_line_polygon_clip    algorithms/cyrus_beck.py:143-241 (polygon: (n_polygon, 2), counter-clockwise, convex,
 *                                      3..32 vertices; all segments against the one polygon)
 *   ct_points_in_polygon               geometry_utils.py:98-147 point_in_polygon (no tolerance), all points against one polygon
 *   ct_points_in_triangles             geometry_utils.py:273-289 (faces: (n_face, n_max_vert) int64, the first three columns
 *                                      are the triangle; face_indices: n int64; an index out of range is CT_ERR_VALUE)
 * ct_profile_binning: diagnostics, device milliseconds of the spatial binning of ct_locate_points alone (points on the device). */
int ct_liang_barsky_line_box_clip(const double *a, const double *b, const double *boxes, int64_t n_boxes, int64_t n,
                                  uint8_t *intersects, double *c, double *d, int32_t mem);
int ct_cohen_sutherland_line_box_clip(const double *a, const double *b, const double *boxes, int64_t n_boxes, int64_t n,
                                      uint8_t *intersects, double *c, double *d, int32_t mem);
int ct_cyrus_beck_line_polygon_clip(const double *a, const double *b, int64_t n, const double *polygon, int32_t n_polygon,
                                    double tolerance, uint8_t *intersects, double *c, double *d, int32_t mem);
int ct_points_in_polygon(const double *points, int64_t n, const double *polygon, int32_t n_polygon, uint8_t *inside, int32_t mem);
int ct_points_in_triangles(const double *points, const int64_t *face_indices, int64_t n, const int64_t *faces, int64_t n_face,
                           int32_t n_max_vert, const double *vertices, int64_t n_vertex, double tolerance, uint8_t *inside,
                           int32_t mem);
int ct_profile_binning(const ct_tree *tree, const double *points, int64_t n, int32_t repeats, double *ms_per_run);
/* Diagnostics behind bench.py's roofline line.  ct_locate_points_stats: what the point traversal touches, summed over the
 * n queries (points on the device): stats[0] node slots read (16 B each), [1] treelet headers read (16 B), [2] cells
 * tested, [3] siblings deferred to the stack, [4] queries started from the entry grid, [5] queries that found a cell.
 * ct_measure_read_bandwidth: read rate of `repeats` sweeps over a buffer of `bytes` bytes (64 MB: the L2 rate; several
 * GB: the HBM rate), GB/s. */
int ct_locate_points_stats(const ct_tree *tree, const double *points, int64_t n, double tolerance, int64_t *stats);
int ct_measure_read_bandwidth(size_t bytes, int32_t repeats, double *gb_per_s);

#ifdef __cplusplus
}
#endif
#endif /* CELLTREE_B200_H */
