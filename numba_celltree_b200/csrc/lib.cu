// lib.cu -- library state (error string, stream, launch counter), the helpers shared by the query
// translation units (scan of per-query counts, order-preserving compaction) and the result / misc entry points.
#include <cub/device/device_scan.cuh>
#include <cub/iterator/transform_input_iterator.cuh>

#include <ctime>
#include <map>
#include <mutex>
#include <thread>
#include <vector>
#include <cstring>
#include <unordered_map>

#include "traverse.cuh"

namespace ct {

// ---- library state -----------------------------------------------------------------------------------
static thread_local std::string g_error;
static thread_local cudaStream_t g_stream = 0;
static int64_t g_launches = 0;

void set_error(const std::string &msg) { g_error = msg; }
cudaStream_t current_stream() { return g_stream; }
void count_launch(int n) { __atomic_fetch_add(&g_launches, (int64_t)n, __ATOMIC_RELAXED); }

static bool g_profile = false;
static PhaseEvents g_events;
static bool g_events_created = false;
static bool g_events_recorded = false;

PhaseEvents *phase_events() {
    if (!g_profile) return nullptr;
    if (!g_events_created) {
        if (cudaEventCreate(&g_events.start) != cudaSuccess || cudaEventCreate(&g_events.ordered) != cudaSuccess ||
            cudaEventCreate(&g_events.done) != cudaSuccess)
            return nullptr;
        g_events_created = true;
    }
    g_events_recorded = true;
    return &g_events;
}

void trace_point(cudaStream_t s, const char *label) {
    static int enabled = -1;
    static double last = 0.0;
    if (enabled < 0) enabled = getenv("CELLTREE_DEBUG") ? 1 : 0;
    if (!enabled) return;
    cudaStreamSynchronize(s);
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    double now = ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
    fprintf(stderr, "[celltree] %-28s +%8.3f ms\n", label, last > 0.0 ? now - last : 0.0);
    last = now;
}

static int g_sort_bits = -1;
int sort_bits_override() {
    static bool env_read = false;
    if (!env_read) {
        env_read = true;
        const char *e = getenv("CELLTREE_SORT_BITS");
        if (e) g_sort_bits = atoi(e);
    }
    return g_sort_bits;
}

// ---- overflow slab of the traversal stacks (common.cuh: DeepStacks, traverse.cuh: Stack) ------------------------------
int DeepScope::init(const ct_tree *tree, int64_t threads, cudaStream_t s) {
    stream = s;
    const int64_t extra = (int64_t)tree->depth - STACK_CAP;
    if (extra <= 0 || threads <= 0) return CT_OK;
    // a column per thread of the launch if that fits the budget, else as many as fit: only threads whose stack really
    // holds more than STACK_CAP deferred siblings take one, and that takes a query crossing that many overlapping nodes
    static int64_t budget = -1;
    if (budget < 0) {
        const char *e = getenv("CELLTREE_DEEP_STACK_MB");
        budget = (e ? atoll(e) : 2048) << 20;
    }
    int64_t slots = budget / (4 * extra);
    if (slots > threads) slots = threads;
    if (slots < 1) slots = 1;
    if (slots > INT32_MAX) slots = INT32_MAX;
    CT_CHECK(slab.alloc((size_t)slots * extra, s));
    CT_CHECK(state.alloc(2, s));
    CT_CUDA(cudaMemsetAsync(state.p, 0, 2 * sizeof(int32_t), s));
    view.slab = slab.p;
    view.state = state.p;
    view.slots = (int32_t)slots;
    return CT_OK;
}
int DeepScope::next_launch() {
    if (view.state) CT_CUDA(cudaMemsetAsync(view.state, 0, sizeof(int32_t), stream));
    return CT_OK;
}
int DeepScope::finish() {
    if (!view.state) return CT_OK;
    int32_t h[2] = {0, 0};
    CT_CUDA(cudaMemcpyAsync(h, view.state, sizeof(h), cudaMemcpyDeviceToHost, stream));
    CT_CUDA(cudaStreamSynchronize(stream));
    if (h[1]) {
        char buf[256];
        snprintf(buf, sizeof(buf), "more than %d queries of one call needed a traversal stack deeper than %d entries at the same "
                 "time: pass fewer queries per call or raise CELLTREE_DEEP_STACK_MB", view.slots, STACK_CAP);
        set_error(buf);
        return CT_ERR_DEPTH;
    }
    return CT_OK;
}

// keep the flagged pairs, order preserved (the NumPy boolean masks of celltree.py:183-184, 226, 268-269)
__global__ void __launch_bounds__(256) k_compact(const int32_t *__restrict__ flag, const int64_t *__restrict__ pos, int64_t n,
                                                 const int32_t *__restrict__ in_i, const int32_t *__restrict__ in_j,
                                                 const double *__restrict__ in_p, int32_t *__restrict__ out_i,
                                                 int32_t *__restrict__ out_j, double *__restrict__ out_p) {
    int64_t k = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (k >= n || !flag[k]) return;
    int64_t o = pos[k];
    out_i[o] = in_i[k];
    out_j[o] = in_j[k];
    if (in_p) out_p[o] = in_p[k];
}

// exclusive scan of int32 counts into int64 offsets[n + 1]; total returned through the last element
int scan_counts(const int32_t *counts, int64_t n, int64_t *offsets, int64_t *total, cudaStream_t s) {
    // offsets[0..n) = exclusive sum; offsets[n] = total.  Scan n + 1 items (counts has a zero sentinel at [n]).
    size_t bytes = 0;
    auto in = cub::TransformInputIterator<int64_t, cub::CastOp<int64_t>, const int32_t *>(counts, cub::CastOp<int64_t>());
    CT_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, offsets, n + 1, s));
    Scratch<char> tmp;
    CT_CHECK(tmp.alloc(bytes, s));
    CT_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, bytes, in, offsets, n + 1, s));
    count_launch(2);
    CT_CUDA(cudaMemcpyAsync(total, offsets + n, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    CT_CUDA(cudaStreamSynchronize(s));
    return CT_OK;
}

// Replace (i, j, payload) of `r` by the flagged subset.
int compact_result(ct_result *r, const int32_t *flag, bool keep_payload, cudaStream_t s) {
    int64_t n = r->size;
    Scratch<int64_t> pos;
    CT_CHECK(pos.alloc(n + 1, s));
    int64_t total = 0;
    // flag has n entries; scan n+1 needs a sentinel: scan n items and add the last flag on the host side
    {
        size_t bytes = 0;
        auto in = cub::TransformInputIterator<int64_t, cub::CastOp<int64_t>, const int32_t *>(flag, cub::CastOp<int64_t>());
        CT_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, pos.p, n, s));
        Scratch<char> tmp;
        CT_CHECK(tmp.alloc(bytes, s));
        if (n > 0) {
            CT_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, bytes, in, pos.p, n, s));
            count_launch(2);
            int64_t last_pos = 0;
            int32_t last_flag = 0;
            CT_CUDA(cudaMemcpyAsync(&last_pos, pos.p + n - 1, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
            CT_CUDA(cudaMemcpyAsync(&last_flag, flag + n - 1, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
            CT_CUDA(cudaStreamSynchronize(s));
            total = last_pos + last_flag;
        }
    }
    trace_point(s, "  compact: scan");
    int32_t *ni = nullptr, *nj = nullptr;
    double *np_ = nullptr;
    CT_CHECK(dalloc(&ni, total, s));
    CT_CHECK(dalloc(&nj, total, s));
    if (keep_payload) CT_CHECK(dalloc(&np_, total, s));
    trace_point(s, "  compact: alloc");
    if (n > 0) {
        k_compact<<<grid_for(n, 256), 256, 0, s>>>(flag, pos.p, n, r->i, r->j, keep_payload ? r->payload : nullptr, ni, nj, np_);
        CT_LAUNCH_CHECK();
    }
    trace_point(s, "  compact: kernel");
    dfree(r->i, s);
    dfree(r->j, s);
    dfree(r->payload, s);
    trace_point(s, "  compact: free");
    r->i = ni;
    r->j = nj;
    r->payload = np_;
    r->width = keep_payload ? 1 : 0;
    r->size = total;
    return CT_OK;
}

}  // namespace ct

using namespace ct;

extern "C" int64_t ct_result_size(const ct_result *r) { return r ? r->size : 0; }
extern "C" int32_t ct_result_payload_width(const ct_result *r) { return r ? r->width : 0; }

extern "C" int ct_result_fetch(const ct_result *r, int64_t *i, int64_t *j, double *payload, int32_t mem) {
    if (!r) {
        set_error("ct_result_fetch: null result");
        return CT_ERR_VALUE;
    }
    cudaStream_t s = current_stream();
    int64_t n = r->size;
    if (n == 0) return CT_OK;
    DevOut<int64_t> oi, oj;
    CT_CHECK(oi.init(i, n, mem, s));
    CT_CHECK(oj.init(j, n, mem, s));
    if (i) {
        CT_CHECK(launch_widen(r->i, n, oi.p, s));
        CT_CHECK(oi.finish(s));
    }
    if (j) {
        CT_CHECK(launch_widen(r->j, n, oj.p, s));
        CT_CHECK(oj.finish(s));
    }
    if (payload && r->width > 0)
        CT_CUDA(cudaMemcpyAsync(payload, r->payload, (size_t)n * r->width * sizeof(double),
                                mem == CT_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, s));
    CT_CUDA(cudaStreamSynchronize(s));
    return CT_OK;
}

extern "C" void ct_result_free(ct_result *r) {
    if (!r) return;
    cudaStream_t s = current_stream();
    dfree(r->i, s);
    dfree(r->j, s);
    dfree(r->payload, s);
    delete r;
}

// ---- caching device memory pool (common.cuh: pool_alloc / pool_free) ---------------------------------------------------
namespace ct {
namespace {
cudaStream_t const STREAM_SYNCED = (cudaStream_t)(intptr_t)-1;  // the block may be used on any stream
struct FreeBlock {
    void *p;
    cudaStream_t stream;  // the stream the block was last used and freed on
    int device;
};
struct DevPool {
    std::mutex lock;
    std::multimap<size_t, FreeBlock> free_blocks;                // by block size
    std::unordered_map<void *, std::pair<size_t, int>> live;      // every block: size, device
    size_t cached_bytes = 0;
    size_t cache_limit = (size_t)64 << 30;
    bool env_read = false;
};
DevPool g_dev_pool;
// sizes: 256 B, powers of two up to 1 MiB, then eight classes per octave (<= 12.5 % slack)
size_t dev_size_class(size_t bytes) {
    if (bytes <= 256) return 256;
    size_t base = 256;
    while (base * 2 <= bytes) base *= 2;
    if (base == bytes) return base;
    if (base < ((size_t)1 << 20)) return base * 2;
    for (int q = 9; q <= 16; q++) {
        size_t c = base / 8 * q;
        if (c >= bytes) return c;
    }
    return base * 2;
}
// cudaFree every cached block of `device` (all devices if < 0); the caller holds the lock
void dev_pool_release_locked(int device) {
    for (auto it = g_dev_pool.free_blocks.begin(); it != g_dev_pool.free_blocks.end();) {
        if (device >= 0 && it->second.device != device) {
            ++it;
            continue;
        }
        int current = 0;
        cudaGetDevice(&current);
        if (current != it->second.device) cudaSetDevice(it->second.device);
        if (it->second.stream != STREAM_SYNCED) cudaStreamSynchronize(it->second.stream);
        cudaFree(it->second.p);
        if (current != it->second.device) cudaSetDevice(current);
        g_dev_pool.cached_bytes -= it->first;
        g_dev_pool.live.erase(it->second.p);
        it = g_dev_pool.free_blocks.erase(it);
    }
}
}  // namespace

int pool_alloc(void **p, size_t bytes, cudaStream_t s) {
    const size_t size = dev_size_class(bytes);
    int device = 0;
    CT_CUDA(cudaGetDevice(&device));
    {
        std::unique_lock<std::mutex> g(g_dev_pool.lock);
        if (!g_dev_pool.env_read) {
            g_dev_pool.env_read = true;
            if (const char *e = getenv("CELLTREE_DEVICE_CACHE_MB")) g_dev_pool.cache_limit = (size_t)atoll(e) << 20;
        }
        // a cached block of this class that was freed on this stream (reuse is in stream order) or on a stream that
        // has been synchronised since; blocks other live streams are still using are left to them
        auto range = g_dev_pool.free_blocks.equal_range(size);
        for (auto it = range.first; it != range.second; ++it) {
            if (it->second.device != device) continue;
            if (it->second.stream == s || it->second.stream == STREAM_SYNCED) {
                *p = it->second.p;
                g_dev_pool.free_blocks.erase(it);
                g_dev_pool.cached_bytes -= size;
                return CT_OK;
            }
        }
    }
    void *q = nullptr;
    cudaError_t e = cudaMalloc(&q, size);
    if (e != cudaSuccess) {  // out of memory: give the cached blocks back and try once more
        cudaGetLastError();
        {
            std::lock_guard<std::mutex> g(g_dev_pool.lock);
            dev_pool_release_locked(device);
        }
        CT_CUDA(cudaMalloc(&q, size));
    }
    std::lock_guard<std::mutex> g(g_dev_pool.lock);
    g_dev_pool.live[q] = {size, device};
    *p = q;
    return CT_OK;
}

void pool_free(void *p, cudaStream_t s) {
    if (!p) return;
    std::unique_lock<std::mutex> g(g_dev_pool.lock);
    auto it = g_dev_pool.live.find(p);
    if (it == g_dev_pool.live.end()) return;
    const size_t size = it->second.first;
    const int device = it->second.second;
    if (g_dev_pool.cached_bytes + size <= g_dev_pool.cache_limit) {
        g_dev_pool.free_blocks.emplace(size, FreeBlock{p, s, device});
        g_dev_pool.cached_bytes += size;
        return;
    }
    g_dev_pool.live.erase(it);
    g.unlock();
    cudaStreamSynchronize(s);
    cudaFree(p);
}

void pool_stream_synced(cudaStream_t s) {
    std::lock_guard<std::mutex> g(g_dev_pool.lock);
    for (auto &kv : g_dev_pool.free_blocks)
        if (kv.second.stream == s) kv.second.stream = STREAM_SYNCED;
}

}  // namespace ct

extern "C" void ct_device_trim(void) {
    std::lock_guard<std::mutex> g(ct::g_dev_pool.lock);
    ct::dev_pool_release_locked(-1);
}

// ---- pinned host memory for results ---------------------------------------------------------------------------
// Device-to-host copies into fresh pageable memory run at a fraction of the PCIe rate (the driver stages them and
// every page is faulted in); results therefore go to page-locked blocks that are recycled through a small cache
// (pinning itself is expensive: it is paid once per size class, not per call).
namespace {
struct HostPool {
    std::mutex lock;
    std::multimap<size_t, void *> free_blocks;     // by block size
    std::unordered_map<void *, size_t> block_size;  // every live block
    size_t cached_bytes = 0;
    size_t cache_limit = (size_t)8 << 30;
};
HostPool g_host_pool;
// size classes 2^k * {1, 1.25, 1.5, 1.75}: a block serves later requests of a similar size (<= 25 % slack)
size_t host_size_class(size_t bytes) {
    size_t base = (size_t)1 << 16;
    while (base * 2 <= bytes) base *= 2;
    for (int q = 4; q <= 8; q++) {
        size_t c = base / 4 * q;
        if (c >= bytes) return c;
    }
    return base * 2;
}
}  // namespace

extern "C" int ct_host_alloc(size_t bytes, void **out) {
    if (!out) {
        set_error("ct_host_alloc: null argument");
        return CT_ERR_VALUE;
    }
    const size_t size = host_size_class(bytes ? bytes : 1);
    {
        std::lock_guard<std::mutex> g(g_host_pool.lock);
        static bool env_read = false;
        if (!env_read) {
            env_read = true;
            if (const char *e = getenv("CELLTREE_PINNED_CACHE_MB")) g_host_pool.cache_limit = (size_t)atoll(e) << 20;
        }
        auto it = g_host_pool.free_blocks.find(size);
        if (it != g_host_pool.free_blocks.end()) {
            *out = it->second;
            g_host_pool.cached_bytes -= size;
            g_host_pool.free_blocks.erase(it);
            return CT_OK;
        }
    }
    void *p = nullptr;
    CT_CUDA(cudaHostAlloc(&p, size, cudaHostAllocPortable));
    std::lock_guard<std::mutex> g(g_host_pool.lock);
    g_host_pool.block_size[p] = size;
    *out = p;
    return CT_OK;
}

extern "C" void ct_host_free(void *p) {
    if (!p) return;
    std::unique_lock<std::mutex> g(g_host_pool.lock);
    auto it = g_host_pool.block_size.find(p);
    if (it == g_host_pool.block_size.end()) return;
    const size_t size = it->second;
    if (g_host_pool.cached_bytes + size <= g_host_pool.cache_limit) {
        g_host_pool.free_blocks.emplace(size, p);
        g_host_pool.cached_bytes += size;
        return;
    }
    g_host_pool.block_size.erase(it);
    g.unlock();
    cudaFreeHost(p);
}

extern "C" void ct_host_trim(void) {
    std::multimap<size_t, void *> blocks;
    {
        std::lock_guard<std::mutex> g(g_host_pool.lock);
        blocks.swap(g_host_pool.free_blocks);
        g_host_pool.cached_bytes = 0;
        for (auto &kv : blocks) g_host_pool.block_size.erase(kv.second);
    }
    for (auto &kv : blocks) cudaFreeHost(kv.second);
}

namespace ct {
int upload_from_host(void *dst_device, const void *src_host, size_t bytes, cudaStream_t s) {
    if (bytes == 0) return CT_OK;
    constexpr size_t STAGED_FROM = (size_t)8 << 20;  // below this the plain copy is as good
    constexpr size_t BLOCK_BYTES = (size_t)32 << 20;
    bool pageable = true;
    {
        cudaPointerAttributes attr;
        if (cudaPointerGetAttributes(&attr, src_host) == cudaSuccess) pageable = attr.type == cudaMemoryTypeUnregistered;
        else cudaGetLastError();
    }
    static int n_threads = -1;
    if (n_threads < 0) {
        const char *e = getenv("CELLTREE_COPY_THREADS");
        int hw = (int)std::thread::hardware_concurrency();
        n_threads = e ? atoi(e) : (hw >= 8 ? 4 : (hw >= 4 ? 2 : 1));  // measured on the pool's 16-core hosts: 4 copies as fast as 8
        if (n_threads < 1) n_threads = 1;
    }
    if (!pageable || bytes < STAGED_FROM || n_threads == 1) {
        CT_CUDA(cudaMemcpyAsync(dst_device, src_host, bytes, cudaMemcpyHostToDevice, s));
        if (pageable) CT_CUDA(cudaStreamSynchronize(s));  // the driver may still be reading a pageable source
        return CT_OK;
    }
    void *stage[2] = {nullptr, nullptr};
    cudaEvent_t drained[2] = {nullptr, nullptr};
    int status = CT_OK;
    auto body = [&]() -> int {
        for (int k = 0; k < 2; k++) {
            CT_CHECK(ct_host_alloc(BLOCK_BYTES, &stage[k]));
            CT_CUDA(cudaEventCreateWithFlags(&drained[k], cudaEventDisableTiming));
        }
        const char *src = static_cast<const char *>(src_host);
        char *dst = static_cast<char *>(dst_device);
        int k = 0;
        bool used[2] = {false, false};
        for (size_t off = 0; off < bytes; off += BLOCK_BYTES, k ^= 1) {
            const size_t len = bytes - off < BLOCK_BYTES ? bytes - off : BLOCK_BYTES;
            if (used[k]) CT_CUDA(cudaEventSynchronize(drained[k]));  // the copy engine is done with this block
            const size_t slice = (len + n_threads - 1) / n_threads;
            std::vector<std::thread> workers;
            for (int t = 1; t < n_threads; t++) {
                const size_t lo = (size_t)t * slice;
                if (lo >= len) break;
                const size_t n = len - lo < slice ? len - lo : slice;
                workers.emplace_back([=]() { memcpy(static_cast<char *>(stage[k]) + lo, src + off + lo, n); });
            }
            memcpy(stage[k], src + off, len < slice ? len : slice);
            for (auto &w : workers) w.join();
            CT_CUDA(cudaMemcpyAsync(dst + off, stage[k], len, cudaMemcpyHostToDevice, s));
            CT_CUDA(cudaEventRecord(drained[k], s));
            used[k] = true;
        }
        for (int q = 0; q < 2; q++)
            if (used[q]) CT_CUDA(cudaEventSynchronize(drained[q]));
        return CT_OK;
    };
    status = body();
    for (int k = 0; k < 2; k++) {
        if (status != CT_OK && drained[k]) cudaEventSynchronize(drained[k]);
        if (drained[k]) cudaEventDestroy(drained[k]);
        ct_host_free(stage[k]);
    }
    return status;
}
}  // namespace ct

extern "C" const char *ct_last_error(void) { return g_error.c_str(); }

extern "C" int ct_device_count(int *count) {
    CT_CUDA(cudaGetDeviceCount(count));
    return CT_OK;
}

extern "C" int ct_set_device(int device) {
    CT_CUDA(cudaSetDevice(device));
    return CT_OK;
}

extern "C" int ct_set_stream(void *cuda_stream) {
    if ((cudaStream_t)cuda_stream != g_stream) {
        // cached blocks freed on the old stream become usable from the new one
        CT_CUDA(cudaStreamSynchronize(g_stream));
        pool_stream_synced(g_stream);
    }
    g_stream = (cudaStream_t)cuda_stream;
    return CT_OK;
}

extern "C" int64_t ct_launch_count(void) { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

extern "C" int ct_set_sort_bits(int32_t bits) {
    (void)sort_bits_override();
    g_sort_bits = bits;
    return CT_OK;
}

extern "C" int ct_profile_enable(int32_t enable) {
    g_profile = enable != 0;
    if (!g_profile) g_events_recorded = false;
    return CT_OK;
}

extern "C" int ct_profile_last(double *order_ms, double *traverse_ms) {
    if (!g_events_created || !g_events_recorded) {
        set_error("ct_profile_last: no profiled ct_locate_points call");
        return CT_ERR_VALUE;
    }
    CT_CUDA(cudaEventSynchronize(g_events.done));
    float a = 0.f, b = 0.f;
    CT_CUDA(cudaEventElapsedTime(&a, g_events.start, g_events.ordered));
    CT_CUDA(cudaEventElapsedTime(&b, g_events.ordered, g_events.done));
    if (order_ms) *order_ms = a;
    if (traverse_ms) *traverse_ms = b;
    return CT_OK;
}
