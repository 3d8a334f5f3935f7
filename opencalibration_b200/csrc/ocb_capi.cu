// C-ABI layer of libocb.so (include/ocb.h): per-thread streams + staging, descriptor residency, launches.
// No CPU fallback anywhere: without a device every compute entry point returns an error.
#include "ocb_internal.cuh"

#include <algorithm>
#include <cstring>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

namespace ocb
{

std::atomic<uint64_t> g_kernel_launches{0};
// device of threads that never called ocb_init / ocb_set_device themselves (OpenMP workers of a one-process-per-GPU
// job must land on the process's GPU, not on device 0): set by the first ocb_init of the process
static std::atomic<int> g_default_device{0};
static std::atomic<bool> g_default_device_set{false};

static thread_local std::string t_last_error;
void set_last_error(const std::string &msg)
{
    t_last_error = msg;
}
int fail_cuda(cudaError_t e, const char *what, const char *file, int line)
{
    char buf[512];
    snprintf(buf, sizeof buf, "CUDA error %d (%s) in %s at %s:%d", (int)e, cudaGetErrorString(e), what, file, line);
    set_last_error(buf);
    cudaGetLastError(); // clear the sticky-less error state
    return -(int)e;
}
int fail_invalid(const char *what)
{
    set_last_error(std::string("invalid argument: ") + what);
    return OCB_E_INVALID;
}

Options &options()
{
    static Options o;
    return o;
}

int sm_count(int device)
{
    static std::mutex mu;
    static std::unordered_map<int, int> cache;
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(device);
    if (it != cache.end())
        return it->second;
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || sms <= 0)
        sms = 148;
    cache[device] = sms;
    return sms;
}

namespace
{

// grow-only buffers owned by one host thread
struct Buf
{
    void *p = nullptr;
    size_t cap = 0;
};

struct ThreadCtx
{
    int device = -1; // -1: follow the process default
    bool ready = false;
    cudaStream_t stream = nullptr; // latency-sensitive per-call work: highest priority
    cudaEvent_t bulk_done = nullptr; // blocking-sync event: a thread waiting for a large submission sleeps instead of
                                     // spinning (the per-pair tail of the previous submission needs the core)
    cudaStream_t bulk = nullptr;   // large batched submissions (ocb_match_pairs): lowest priority, so that the small
                                   // kernels of other host threads (RANSAC scoring of the previous submission) are
                                   // dispatched ahead of the thousands of queued matching CTAs
    Buf dev, pinned;
    // correspondences bound to this thread by ocb_corr_bind (one RANSAC run scores many model batches against the
    // same set): [n][7] as given, [n] double4 in index order, [n] double4 + positions in evaluation order
    Buf bound;
    // batch binding (ocb_corr_bind_batch): concatenated [n][7] rows and evaluation orders of many sets
    struct BatchSet
    {
        size_t o_c7 = 0, o_ord = 0, n = 0;
        bool has_order = false;
    };
    Buf batch;
    std::vector<BatchSet> batch_sets;
    bool batch_valid = false;
    size_t bound_n = 0;
    bool bound_valid = false, bound_has_order = false;
    int ready_device = -1;

    int ensure()
    {
        if (device < 0)
            device = g_default_device.load();
        if (ready && ready_device == device)
        {
            OCB_CUDA(cudaSetDevice(device));
            return 0;
        }
        release();
        int n = 0;
        cudaError_t e = cudaGetDeviceCount(&n);
        if (e != cudaSuccess || n == 0)
        {
            set_last_error("no CUDA device available (libocb has no CPU fallback)");
            cudaGetLastError();
            return OCB_E_NO_DEVICE;
        }
        if (device < 0 || device >= n)
            return fail_invalid("device index out of range");
        OCB_CUDA(cudaSetDevice(device));
        int prio_low = 0, prio_high = 0;
        OCB_CUDA(cudaDeviceGetStreamPriorityRange(&prio_low, &prio_high));
        OCB_CUDA(cudaStreamCreateWithPriority(&stream, cudaStreamNonBlocking, prio_high));
        OCB_CUDA(cudaStreamCreateWithPriority(&bulk, cudaStreamNonBlocking, prio_low));
        OCB_CUDA(cudaEventCreateWithFlags(&bulk_done, cudaEventBlockingSync | cudaEventDisableTiming));
        ready = true;
        ready_device = device;
        return 0;
    }
    bool blocking_sync = false; // ocb_set_thread_blocking_sync
    // waits for everything enqueued on the per-call stream: spinning (lowest latency) or, when the thread asked for it,
    // sleeping on the blocking-sync event
    int wait_stream()
    {
        if (!blocking_sync)
        {
            OCB_CUDA(cudaStreamSynchronize(stream));
            return 0;
        }
        OCB_CUDA(cudaEventRecord(bulk_done, stream));
        OCB_CUDA(cudaEventSynchronize(bulk_done));
        return 0;
    }
    // waits for everything enqueued on the bulk stream without spinning
    int wait_bulk()
    {
        OCB_CUDA(cudaEventRecord(bulk_done, bulk));
        OCB_CUDA(cudaEventSynchronize(bulk_done));
        return 0;
    }
    int dev_reserve(size_t bytes)
    {
        if (bytes <= dev.cap)
            return 0;
        if (dev.p)
        {
            OCB_CUDA(cudaStreamSynchronize(stream));
            OCB_CUDA(cudaFree(dev.p));
            dev.p = nullptr;
            dev.cap = 0;
        }
        const size_t cap = std::max(bytes + bytes / 4, (size_t)1 << 20);
        OCB_CUDA(cudaMalloc(&dev.p, cap));
        dev.cap = cap;
        return 0;
    }
    int pinned_reserve(size_t bytes)
    {
        if (bytes <= pinned.cap)
            return 0;
        if (pinned.p)
        {
            OCB_CUDA(cudaStreamSynchronize(stream));
            OCB_CUDA(cudaFreeHost(pinned.p));
            pinned.p = nullptr;
            pinned.cap = 0;
        }
        const size_t cap = std::max(bytes + bytes / 4, (size_t)1 << 20);
        OCB_CUDA(cudaHostAlloc(&pinned.p, cap, cudaHostAllocDefault));
        pinned.cap = cap;
        return 0;
    }
    void release()
    {
        if (ready)
        {
            cudaSetDevice(ready_device);
            if (stream)
                cudaStreamSynchronize(stream);
            if (dev.p)
                cudaFree(dev.p);
            if (pinned.p)
                cudaFreeHost(pinned.p);
            if (bound.p)
                cudaFree(bound.p);
            if (batch.p)
                cudaFree(batch.p);
            if (stream)
                cudaStreamDestroy(stream);
            if (bulk)
            {
                cudaStreamSynchronize(bulk);
                cudaStreamDestroy(bulk);
            }
            if (bulk_done)
                cudaEventDestroy(bulk_done);
        }
        dev = Buf();
        pinned = Buf();
        bound = Buf();
        bound_valid = false;
        batch = Buf();
        batch_valid = false;
        batch_sets.clear();
        stream = nullptr;
        bulk = nullptr;
        bulk_done = nullptr;
        ready = false;
        ready_device = -1;
    }
    ~ThreadCtx()
    {
        // worker threads come and go (std::async, OpenMP teams): give their streams and buffers back. At process
        // exit the runtime may already be unloading; the calls then fail harmlessly and the OS reclaims the rest.
        release();
        cudaGetLastError();
    }
};
thread_local ThreadCtx t_ctx;

inline size_t align_up(size_t v, size_t a)
{
    return (v + a - 1) / a * a;
}

// carve regions out of one buffer, 256-byte aligned
struct Carver
{
    size_t off = 0;
    size_t take(size_t bytes)
    {
        const size_t o = off;
        off = align_up(off + bytes, 256);
        return o;
    }
};

// Is `p` page-locked host memory the copy engines can read directly?
bool is_pinned_host(const void *p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess)
    {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}

// Host -> device through the thread's pinned staging area unless the source is already pinned.
int upload(ThreadCtx &c, void *d_dst, const void *h_src, size_t bytes, size_t staging_off)
{
    if (bytes == 0)
        return 0;
    if (is_pinned_host(h_src))
    {
        OCB_CUDA(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, c.stream));
    }
    else
    {
        memcpy(static_cast<char *>(c.pinned.p) + staging_off, h_src, bytes);
        OCB_CUDA(cudaMemcpyAsync(d_dst, static_cast<char *>(c.pinned.p) + staging_off, bytes, cudaMemcpyHostToDevice,
                                 c.stream));
    }
    return 0;
}

// Descriptor sets come and go with every LinkStage batch; cudaMalloc / cudaFree of hundreds of MB cost tens to hundreds
// of milliseconds (and cudaFree synchronises the device), so set storage comes from the device's stream-ordered
// memory pool with a release threshold that keeps freed blocks cached for the next batch.
cudaError_t set_storage_alloc(void **p, size_t bytes, int device)
{
    static std::mutex mu;
    static std::unordered_map<int, bool> configured;
    {
        std::lock_guard<std::mutex> lk(mu);
        if (!configured[device])
        {
            cudaMemPool_t pool;
            if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess)
            {
                uint64_t keep = ~0ull;
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
            }
            cudaGetLastError();
            configured[device] = true;
        }
    }
    cudaError_t e = cudaMallocAsync(p, bytes, cudaStreamPerThread);
    if (e == cudaSuccess)
        e = cudaStreamSynchronize(cudaStreamPerThread);
    return e;
}
void set_storage_free(void *p)
{
    if (p)
        cudaFreeAsync(p, cudaStreamPerThread);
}

// One device allocation shared by the sets of a batched registration; freed with its last set.
struct Arena
{
    void *base = nullptr;
    int device = 0;
    size_t live = 0;
};
struct DescSet
{
    void *d_rows = nullptr;
    size_t n = 0;
    int device = 0;
    Arena *arena = nullptr; // nullptr: d_rows is its own allocation
    // keypoint locations + camera model (ocb_register_images_batch): what K6 needs after the match
    void *d_xy = nullptr; // [n] double2, inside the same arena
    ocb_camera camera{};
};
// caller holds g_sets_mu
void free_set_storage(DescSet &s)
{
    int cur = 0;
    cudaGetDevice(&cur);
    if (s.arena)
    {
        if (--s.arena->live == 0)
        {
            cudaSetDevice(s.arena->device);
            set_storage_free(s.arena->base);
            delete s.arena;
        }
    }
    else if (s.d_rows)
    {
        cudaSetDevice(s.device);
        set_storage_free(s.d_rows);
    }
    cudaSetDevice(cur);
    s.d_rows = nullptr;
    s.d_xy = nullptr;
    s.arena = nullptr;
}
std::mutex g_sets_mu;
std::unordered_map<uint64_t, DescSet> g_sets; // key = set id; one device per id

bool aligned16(const void *p)
{
    return (reinterpret_cast<uintptr_t>(p) & 15u) == 0;
}

// An evaluation order indexes the correspondences on the device (score_models.cu: c7 + order[p] * 7); an entry out of
// range would be a wild device read, so it is rejected on the host like sample and list indices are.
bool order_in_range(const uint32_t *order, size_t n)
{
    if (!order)
        return true;
    uint32_t worst = 0;
    for (size_t p = 0; p < n; p++)
        worst = std::max(worst, order[p]);
    return n == 0 || worst < n;
}

} // namespace
} // namespace ocb

using namespace ocb;

extern "C"
{

    int ocb_device_count(void)
    {
        int n = 0;
        if (cudaGetDeviceCount(&n) != cudaSuccess)
        {
            cudaGetLastError();
            return 0;
        }
        return n;
    }

    int ocb_init(int device)
    {
        t_ctx.device = device;
        const int rc = t_ctx.ensure();
        if (rc == 0 && !g_default_device_set.exchange(true))
            g_default_device.store(device);
        return rc;
    }

    int ocb_set_device(int device)
    {
        t_ctx.device = device;
        return 0;
    }

    int ocb_set_thread_blocking_sync(int on)
    {
        t_ctx.blocking_sync = on != 0;
        return 0;
    }

    int ocb_current_device(void)
    {
        return t_ctx.device >= 0 ? t_ctx.device : g_default_device.load();
    }

    void ocb_shutdown(void)
    {
        {
            std::lock_guard<std::mutex> lk(g_sets_mu);
            for (auto &kv : g_sets)
                free_set_storage(kv.second);
            g_sets.clear();
        }
        t_ctx.release();
    }

    const char *ocb_last_error(void)
    {
        return t_last_error.c_str();
    }

    const char *ocb_version(void)
    {
        return "opencalibration_b200 0.1 (sm_100a)";
    }

    uint64_t ocb_kernel_launches(void)
    {
        return g_kernel_launches.load(std::memory_order_relaxed);
    }

    int ocb_set_option(const char *key, int64_t value)
    {
        if (!key)
            return fail_invalid("key");
        Options &o = options();
        if (!strcmp(key, "k1_variant"))
            o.k1_variant = (int)value;
        else if (!strcmp(key, "k1_items_per_sm"))
            o.k1_items_per_sm = (int)value;
        else if (!strcmp(key, "k2_variant"))
            o.k2_variant = (int)value;
        else if (!strcmp(key, "k2_hg"))
            o.k2_hg = (int)value;
        else if (!strcmp(key, "k1_update"))
            o.k1_update = (int)value;
        else if (!strcmp(key, "k1_bf_rows"))
            o.k1_bf_rows = (int)value;
        else if (!strcmp(key, "k1_engine"))
            o.k1_engine = (int)value;
        else if (!strcmp(key, "k1t_variant"))
            o.k1t_variant = (int)value;
        else
            return fail_invalid("unknown option");
        return 0;
    }

    int64_t ocb_get_option(const char *key)
    {
        if (!key)
            return OCB_E_INVALID;
        Options &o = options();
        if (!strcmp(key, "k1_variant"))
            return o.k1_variant;
        if (!strcmp(key, "k1_items_per_sm"))
            return o.k1_items_per_sm;
        if (!strcmp(key, "k2_variant"))
            return o.k2_variant;
        if (!strcmp(key, "k2_hg"))
            return o.k2_hg;
        if (!strcmp(key, "k1_update"))
            return o.k1_update;
        if (!strcmp(key, "k1_bf_rows"))
            return o.k1_bf_rows;
        if (!strcmp(key, "k1_engine"))
            return o.k1_engine;
        if (!strcmp(key, "k1t_variant"))
            return o.k1t_variant;
        if (!strcmp(key, "k1_queries_per_cta"))
            return k1_queries_per_cta();
        return OCB_E_INVALID;
    }

    // ---------------------------------------------------------------------------------------------------
    // K1
    // ---------------------------------------------------------------------------------------------------
    size_t ocb_match_top2_workspace_bytes(size_t n1, size_t n2, int with_col_best)
    {
        // worst case of the planner: every candidate tile its own split
        K1Problem pr;
        memset(&pr, 0, sizeof pr);
        pr.n_q = (uint32_t)n1, pr.n_c = (uint32_t)n2;
        pr.col_out = with_col_best ? reinterpret_cast<uint32_t *>(16) : nullptr;
        k1_plan(&pr, 1, 148, 1 << 20);
        // whichever engine the call ends up using must fit
        return std::max(align_up(k1_state_bytes(pr), 256) + 256, k1t_workspace_bytes(n1, n2, with_col_best != 0));
    }

    int ocb_match_top2_device(const void *d_q, size_t n1, const void *d_c, size_t n2, void *d_out,
                              void *d_col_best_q, void *d_workspace, size_t workspace_bytes, void *stream)
    {
        if (n1 >= 0xFFFFFFFFull || n2 >= 0xFFFFFFFFull)
            return fail_invalid("n1/n2 must fit in 32 bits");
        if ((n1 && (!d_q || !d_out)) || (n2 && !d_c))
            return fail_invalid("null device pointer");
        if (!aligned16(d_q) || !aligned16(d_c) || (reinterpret_cast<uintptr_t>(d_out) & 7u) ||
            (reinterpret_cast<uintptr_t>(d_workspace) & 255u))
            return fail_invalid("device pointers must be 16-byte aligned (workspace 256)");
        int dev = 0;
        OCB_CUDA(cudaGetDevice(&dev));
        cudaStream_t st = static_cast<cudaStream_t>(stream);
        const bool col = d_col_best_q != nullptr && n2 > 0;
        if (col && n1 == 0) // no query: every column is empty and no CTA runs
        {
            OCB_CUDA(cudaMemsetAsync(d_col_best_q, 0xFF, n2 * sizeof(uint32_t), st));
            return 0;
        }
        // Engine: a pair large enough to fill the tensor pipe goes to K1T (the same records from an exact s8
        // contraction, hamming_tensor.cu), everything else -- small pairs, pairs beyond K1T's position range -- to K1.
        const int engine = options().k1_engine;
        const bool big = n1 >= 512 && n2 >= 512;
        if (n1 && n2 && k1t_supports(n1, n2) && (engine == 2 || (engine == 0 && big)) &&
            k1t_workspace_bytes(n1, n2, col) <= workspace_bytes)
            return k1t_launch(d_q, n1, d_c, n2, static_cast<ocb_top2 *>(d_out),
                              col ? static_cast<uint32_t *>(d_col_best_q) : nullptr, d_workspace, sm_count(dev), st);
        K1Problem pr;
        memset(&pr, 0, sizeof pr);
        pr.q = static_cast<const uint4 *>(d_q), pr.n_q = (uint32_t)n1;
        pr.c = static_cast<const uint4 *>(d_c), pr.n_c = (uint32_t)n2;
        pr.out = static_cast<ocb_top2 *>(d_out);
        pr.col_out = col ? static_cast<uint32_t *>(d_col_best_q) : nullptr;
        K1Plan plan = k1_plan(&pr, 1, sm_count(dev));
        // workspace = the kernel's state block (tickets, merge state, column keys), zeroed every call
        const size_t zero_bytes = k1_state_bytes(pr);
        if (zero_bytes > workspace_bytes)
            return fail_invalid("workspace too small");
        k1_bind_state(pr, d_workspace);
        if (zero_bytes)
            OCB_CUDA(cudaMemsetAsync(d_workspace, 0, zero_bytes, st));
        return k1_launch(nullptr, &pr, 1, plan, st);
    }

    int ocb_match_top2(const uint64_t *q, size_t n1, const uint64_t *c, size_t n2, ocb_top2 *out, uint32_t *col_best_q)
    {
        if (n1 >= 0xFFFFFFFFull || n2 >= 0xFFFFFFFFull)
            return fail_invalid("n1/n2 must fit in 32 bits");
        if ((n1 && (!q || !out)) || (n2 && !c))
            return fail_invalid("null pointer");
        ThreadCtx &cx = t_ctx;
        int rc = cx.ensure();
        if (rc)
            return rc;
        if (n1 == 0 && (n2 == 0 || !col_best_q))
            return 0;
        const bool col = col_best_q != nullptr;
        const size_t qb = n1 * OCB_ROW_BYTES, cb = n2 * OCB_ROW_BYTES;
        const size_t ws_bytes = ocb_match_top2_workspace_bytes(n1, n2, col);
        Carver cv;
        const size_t o_q = cv.take(qb), o_c = cv.take(cb), o_out = cv.take(n1 * sizeof(ocb_top2));
        const size_t o_col = cv.take(col ? n2 * sizeof(uint32_t) : 0);
        const size_t o_ws = cv.take(ws_bytes);
        rc = cx.dev_reserve(cv.off);
        if (rc)
            return rc;
        // staging: rows in (when the caller's memory is pageable) and results out
        Carver sv;
        const size_t s_q = sv.take(qb), s_c = sv.take(cb), s_out = sv.take(n1 * sizeof(ocb_top2));
        const size_t s_col = sv.take(col ? n2 * sizeof(uint32_t) : 0);
        rc = cx.pinned_reserve(sv.off);
        if (rc)
            return rc;
        char *d = static_cast<char *>(cx.dev.p);
        if ((rc = upload(cx, d + o_q, q, qb, s_q)))
            return rc;
        if ((rc = upload(cx, d + o_c, c, cb, s_c)))
            return rc;
        rc = ocb_match_top2_device(d + o_q, n1, d + o_c, n2, d + o_out, col ? d + o_col : nullptr, d + o_ws, ws_bytes,
                                   cx.stream);
        if (rc)
            return rc;
        char *hp = static_cast<char *>(cx.pinned.p);
        const bool out_pinned = is_pinned_host(out);
        if (n1)
            OCB_CUDA(cudaMemcpyAsync(out_pinned ? (void *)out : (void *)(hp + s_out), d + o_out, n1 * sizeof(ocb_top2),
                                     cudaMemcpyDeviceToHost, cx.stream));
        const bool col_pinned = col && is_pinned_host(col_best_q);
        if (col && n2)
            OCB_CUDA(cudaMemcpyAsync(col_pinned ? (void *)col_best_q : (void *)(hp + s_col), d + o_col,
                                     n2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, cx.stream));
        OCB_CUDA(cudaStreamSynchronize(cx.stream));
        if (n1 && !out_pinned)
            memcpy(out, hp + s_out, n1 * sizeof(ocb_top2));
        if (col && n2 && !col_pinned)
            memcpy(col_best_q, hp + s_col, n2 * sizeof(uint32_t));
        return 0;
    }

    // Gathers rows base[idx[k] * stride .. +64) (idx == NULL: k * stride) into dst.
    static void gather_rows(char *dst, const void *base, size_t stride, const size_t *idx, size_t n, size_t first = 0,
                            size_t elem = OCB_ROW_BYTES)
    {
        const char *b = static_cast<const char *>(base) + first * stride;
        if (!idx && stride == elem)
        {
            memcpy(dst, b, n * elem);
            return;
        }
        for (size_t k = 0; k < n; k++)
            memcpy(dst + k * elem, b + (idx ? idx[k] : k) * stride, elem);
    }

    int ocb_match_top2_strided(const void *rows1, size_t stride1, const size_t *idx1, size_t n1, const void *rows2,
                               size_t stride2, const size_t *idx2, size_t n2, ocb_top2 *out, uint32_t *col_best_q)
    {
        if (n1 >= 0xFFFFFFFFull || n2 >= 0xFFFFFFFFull)
            return fail_invalid("n1/n2 must fit in 32 bits");
        if ((n1 && (!rows1 || !out)) || (n2 && !rows2))
            return fail_invalid("null pointer");
        if (stride1 < OCB_ROW_BYTES || stride2 < OCB_ROW_BYTES)
            return fail_invalid("row stride below 64 bytes");
        ThreadCtx &cx = t_ctx;
        int rc = cx.ensure();
        if (rc)
            return rc;
        if (n1 == 0 && (n2 == 0 || !col_best_q))
            return 0;
        const bool col = col_best_q != nullptr;
        const size_t qb = n1 * OCB_ROW_BYTES, cb = n2 * OCB_ROW_BYTES;
        const size_t ws_bytes = ocb_match_top2_workspace_bytes(n1, n2, col);
        Carver cv;
        const size_t o_q = cv.take(qb), o_c = cv.take(cb), o_out = cv.take(n1 * sizeof(ocb_top2));
        const size_t o_col = cv.take(col ? n2 * sizeof(uint32_t) : 0);
        const size_t o_ws = cv.take(ws_bytes);
        rc = cx.dev_reserve(cv.off);
        if (rc)
            return rc;
        Carver sv;
        const size_t s_q = sv.take(qb), s_c = sv.take(cb), s_out = sv.take(n1 * sizeof(ocb_top2));
        const size_t s_col = sv.take(col ? n2 * sizeof(uint32_t) : 0);
        rc = cx.pinned_reserve(sv.off);
        if (rc)
            return rc;
        char *d = static_cast<char *>(cx.dev.p);
        char *hp = static_cast<char *>(cx.pinned.p);
        // the candidate copy is in flight while the query rows are gathered
        if (n2)
        {
            gather_rows(hp + s_c, rows2, stride2, idx2, n2);
            OCB_CUDA(cudaMemcpyAsync(d + o_c, hp + s_c, cb, cudaMemcpyHostToDevice, cx.stream));
        }
        if (n1)
        {
            gather_rows(hp + s_q, rows1, stride1, idx1, n1);
            OCB_CUDA(cudaMemcpyAsync(d + o_q, hp + s_q, qb, cudaMemcpyHostToDevice, cx.stream));
        }
        rc = ocb_match_top2_device(d + o_q, n1, d + o_c, n2, d + o_out, col ? d + o_col : nullptr, d + o_ws, ws_bytes,
                                   cx.stream);
        if (rc)
            return rc;
        if (n1)
            OCB_CUDA(cudaMemcpyAsync(hp + s_out, d + o_out, n1 * sizeof(ocb_top2), cudaMemcpyDeviceToHost, cx.stream));
        if (col && n2)
            OCB_CUDA(cudaMemcpyAsync(hp + s_col, d + o_col, n2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, cx.stream));
        OCB_CUDA(cudaStreamSynchronize(cx.stream));
        if (n1)
            memcpy(out, hp + s_out, n1 * sizeof(ocb_top2));
        if (col && n2)
            memcpy(col_best_q, hp + s_col, n2 * sizeof(uint32_t));
        return 0;
    }

    // ---------------------------------------------------------------------------------------------------
    // descriptor residency + batched pairs
    // ---------------------------------------------------------------------------------------------------
    int ocb_register_descriptors(uint64_t set_id, const uint64_t *rows, size_t n)
    {
        if (n >= 0xFFFFFFFFull)
            return fail_invalid("n must fit in 32 bits");
        if (n && !rows)
            return fail_invalid("rows");
        ThreadCtx &cx = t_ctx;
        int rc = cx.ensure();
        if (rc)
            return rc;
        DescSet s;
        s.n = n;
        s.device = cx.device;
        if (n)
        {
            OCB_CUDA(set_storage_alloc(&s.d_rows, n * OCB_ROW_BYTES, cx.device));
            cudaError_t e = cudaMemcpyAsync(s.d_rows, rows, n * OCB_ROW_BYTES, cudaMemcpyHostToDevice, cx.stream);
            if (e == cudaSuccess)
                e = cudaStreamSynchronize(cx.stream);
            if (e != cudaSuccess)
            {
                set_storage_free(s.d_rows);
                return fail_cuda(e, "upload descriptor set", __FILE__, __LINE__);
            }
        }
        std::lock_guard<std::mutex> lk(g_sets_mu);
        auto it = g_sets.find(set_id);
        if (it != g_sets.end())
            free_set_storage(it->second);
        g_sets[set_id] = s;
        return 0;
    }

    int ocb_register_images_batch(const ocb_image_source *sources, size_t count)
    {
        if (count == 0)
            return 0;
        if (!sources)
            return fail_invalid("sources");
        ThreadCtx &cx = t_ctx;
        int rc = cx.ensure();
        if (rc)
            return rc;
        constexpr size_t XY_BYTES = 2 * sizeof(double);
        size_t total_rows = 0, total_xy = 0;
        for (size_t i = 0; i < count; i++)
        {
            if (sources[i].n >= 0xFFFFFFFFull)
                return fail_invalid("n must fit in 32 bits");
            if (sources[i].n && (!sources[i].rows || sources[i].stride < OCB_ROW_BYTES))
                return fail_invalid("rows / stride");
            if (sources[i].n && sources[i].xy && sources[i].xy_stride < XY_BYTES)
                return fail_invalid("xy stride below 16 bytes");
            total_rows += sources[i].n;
            total_xy += sources[i].xy ? sources[i].n : 0;
        }
        {
            // the same id twice in one batch would release the first entry's share of the arena while the batch is
            // still being entered
            std::vector<uint64_t> ids(count);
            for (size_t i = 0; i < count; i++)
                ids[i] = sources[i].set_id;
            std::sort(ids.begin(), ids.end());
            if (std::adjacent_find(ids.begin(), ids.end()) != ids.end())
                return fail_invalid("duplicate set_id in one batch");
        }
        Arena *arena = nullptr;
        char *d_base = nullptr;
        const size_t xy_base = total_rows * OCB_ROW_BYTES; // keypoints follow the rows inside the allocation
        if (total_rows)
        {
            void *p = nullptr;
            OCB_CUDA(set_storage_alloc(&p, xy_base + total_xy * XY_BYTES, cx.device));
            arena = new Arena;
            arena->base = p, arena->device = cx.device, arena->live = 0;
            d_base = static_cast<char *>(p);
        }
        // gather -> page-locked staging -> device, double-buffered so that the copy engine works while the next
        // block is gathered; first every set's rows, then every set's keypoints
        const size_t block_bytes = (size_t)4 << 20;
        rc = cx.pinned_reserve(2 * block_bytes);
        cudaEvent_t ev[2] = {nullptr, nullptr};
        if (!rc && total_rows)
        {
            char *hp = static_cast<char *>(cx.pinned.p);
            cudaError_t e = cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming);
            if (e == cudaSuccess)
                e = cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming);
            size_t blk = 0;
            for (int pass = 0; pass < 2 && e == cudaSuccess; pass++)
            {
                const size_t elem = pass == 0 ? (size_t)OCB_ROW_BYTES : XY_BYTES;
                const size_t block_elems = block_bytes / elem;
                size_t done = 0;
                char *d_dst = d_base + (pass == 0 ? 0 : xy_base);
                for (size_t i = 0; i < count && e == cudaSuccess; i++)
                {
                    const void *base = pass == 0 ? sources[i].rows : sources[i].xy;
                    const size_t stride = pass == 0 ? sources[i].stride : sources[i].xy_stride;
                    if (!base)
                        continue;
                    size_t k = 0;
                    while (k < sources[i].n && e == cudaSuccess)
                    {
                        const size_t take = std::min(block_elems, sources[i].n - k);
                        const int b = (int)(blk & 1);
                        if (blk >= 2)
                            e = cudaEventSynchronize(ev[b]);
                        if (e != cudaSuccess)
                            break;
                        gather_rows(hp + (size_t)b * block_bytes, base, stride,
                                    sources[i].idx ? sources[i].idx + k : nullptr, take, sources[i].idx ? 0 : k, elem);
                        e = cudaMemcpyAsync(d_dst + done * elem, hp + (size_t)b * block_bytes, take * elem,
                                            cudaMemcpyHostToDevice, cx.stream);
                        if (e == cudaSuccess)
                            e = cudaEventRecord(ev[b], cx.stream);
                        k += take, done += take, blk++;
                    }
                }
            }
            if (e == cudaSuccess)
                e = cudaStreamSynchronize(cx.stream);
            for (cudaEvent_t v : ev)
                if (v)
                    cudaEventDestroy(v);
            if (e != cudaSuccess)
                rc = fail_cuda(e, "upload descriptor sets", __FILE__, __LINE__);
        }
        if (rc)
        {
            if (arena)
            {
                set_storage_free(arena->base);
                delete arena;
            }
            return rc;
        }
        std::lock_guard<std::mutex> lk(g_sets_mu);
        size_t off = 0, off_xy = 0;
        for (size_t i = 0; i < count; i++)
        {
            auto it = g_sets.find(sources[i].set_id);
            if (it != g_sets.end())
                free_set_storage(it->second);
            DescSet s;
            s.n = sources[i].n;
            s.device = cx.device;
            s.camera = sources[i].camera;
            if (sources[i].n)
            {
                s.d_rows = d_base + off * OCB_ROW_BYTES;
                if (sources[i].xy)
                {
                    s.d_xy = d_base + xy_base + off_xy * XY_BYTES;
                    off_xy += sources[i].n;
                }
                s.arena = arena;
                arena->live++;
            }
            off += sources[i].n;
            g_sets[sources[i].set_id] = s;
        }
        return 0;
    }

    int ocb_register_descriptors_batch(const ocb_set_source *sources, size_t count)
    {
        if (count == 0)
            return 0;
        if (!sources)
            return fail_invalid("sources");
        std::vector<ocb_image_source> img(count);
        for (size_t i = 0; i < count; i++)
        {
            memset(&img[i], 0, sizeof img[i]);
            img[i].set_id = sources[i].set_id, img[i].rows = sources[i].rows, img[i].stride = sources[i].stride;
            img[i].idx = sources[i].idx, img[i].n = sources[i].n;
        }
        return ocb_register_images_batch(img.data(), count);
    }

    void *ocb_host_alloc(size_t bytes)
    {
        void *p = nullptr;
        if (t_ctx.ensure() != 0)
            return nullptr;
        if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess)
        {
            cudaGetLastError();
            set_last_error("cudaHostAlloc failed");
            return nullptr;
        }
        return p;
    }

    void ocb_host_free(void *p)
    {
        if (p)
            cudaFreeHost(p);
    }

    int ocb_unregister_descriptors(uint64_t set_id)
    {
        std::lock_guard<std::mutex> lk(g_sets_mu);
        auto it = g_sets.find(set_id);
        if (it == g_sets.end())
        {
            set_last_error("unknown descriptor set");
            return OCB_E_NOT_FOUND;
        }
        free_set_storage(it->second);
        g_sets.erase(it);
        return 0;
    }

    int ocb_match_pairs(const ocb_pair *pairs, size_t n_pairs, ocb_top2 *out, const uint64_t *out_offsets)
    {
        if (n_pairs == 0)
            return 0;
        if (!pairs || !out || !out_offsets)
            return fail_invalid("null pointer");
        if (n_pairs >= (1ull << 31))
            return fail_invalid("too many pairs");
        ThreadCtx &cx = t_ctx;
        int rc = cx.ensure();
        if (rc)
            return rc;
        cudaStream_t bulk = cx.bulk;
        std::vector<K1Problem> pr(n_pairs);
        memset(pr.data(), 0, sizeof(K1Problem) * n_pairs);
        uint64_t out_end = 0;
        {
            std::lock_guard<std::mutex> lk(g_sets_mu);
            for (size_t p = 0; p < n_pairs; p++)
            {
                auto a = g_sets.find(pairs[p].query_set), b = g_sets.find(pairs[p].candidate_set);
                if (a == g_sets.end() || b == g_sets.end())
                {
                    set_last_error("unknown descriptor set in pair list");
                    return OCB_E_NOT_FOUND;
                }
                if (a->second.device != cx.device || b->second.device != cx.device)
                    return fail_invalid("descriptor set registered on another device");
                pr[p].q = static_cast<const uint4 *>(a->second.d_rows), pr[p].n_q = (uint32_t)a->second.n;
                pr[p].c = static_cast<const uint4 *>(b->second.d_rows), pr[p].n_c = (uint32_t)b->second.n;
                out_end = std::max(out_end, out_offsets[p] + a->second.n);
            }
        }
        K1Plan plan = k1_plan(pr.data(), n_pairs, sm_count(cx.device));
        if ((uint64_t)plan.total_items >= 0x7FFFFFFFull)
            return fail_invalid("too many work items for one submission");
        const size_t table_bytes = sizeof(K1Problem) * n_pairs;
        size_t state_total = 0;
        for (size_t p = 0; p < n_pairs; p++)
            state_total += align_up(k1_state_bytes(pr[p]), 16);
        Carver cv;
        const size_t o_tab = cv.take(table_bytes), o_out = cv.take(out_end * sizeof(ocb_top2));
        const size_t o_state = cv.take(state_total);
        rc = cx.dev_reserve(cv.off);
        if (rc)
            return rc;
        Carver sv;
        const size_t s_tab = sv.take(table_bytes), s_out = sv.take(out_end * sizeof(ocb_top2));
        rc = cx.pinned_reserve(sv.off);
        if (rc)
            return rc;
        char *d = static_cast<char *>(cx.dev.p);
        char *hp = static_cast<char *>(cx.pinned.p);
        size_t state_off = 0;
        for (size_t p = 0; p < n_pairs; p++)
        {
            pr[p].out = reinterpret_cast<ocb_top2 *>(d + o_out) + out_offsets[p];
            k1_bind_state(pr[p], d + o_state + state_off);
            state_off += align_up(k1_state_bytes(pr[p]), 16);
        }
        if (state_total)
            OCB_CUDA(cudaMemsetAsync(d + o_state, 0, state_total, bulk));
        const K1Problem *d_tab = nullptr;
        if (n_pairs > (size_t)K1_INLINE)
        {
            memcpy(hp + s_tab, pr.data(), table_bytes);
            OCB_CUDA(cudaMemcpyAsync(d + o_tab, hp + s_tab, table_bytes, cudaMemcpyHostToDevice, bulk));
            d_tab = reinterpret_cast<const K1Problem *>(d + o_tab);
        }
        rc = k1_launch(d_tab, pr.data(), n_pairs, plan, bulk);
        if (rc)
            return rc;
        const bool out_pinned = is_pinned_host(out);
        if (out_end)
            OCB_CUDA(cudaMemcpyAsync(out_pinned ? (void *)out : (void *)(hp + s_out), d + o_out,
                                     out_end * sizeof(ocb_top2), cudaMemcpyDeviceToHost, bulk));
        if ((rc = cx.wait_bulk()))
            return rc;
        if (out_end && !out_pinned)
            memcpy(out, hp + s_out, out_end * sizeof(ocb_top2));
        return 0;
    }

    static int match_pairs_tail(const ocb_pair *pairs, size_t n_pairs, ocb_match *out, size_t out_capacity,
                                uint64_t *out_offsets, bool sorted, uint32_t *quality_order);
    int ocb_match_pairs_ratio(const ocb_pair *pairs, size_t n_pairs, ocb_match *out, size_t out_capacity,
                              uint64_t *out_offsets)
    {
        return match_pairs_tail(pairs, n_pairs, out, out_capacity, out_offsets, false, nullptr);
    }
    int ocb_match_pairs_sorted(const ocb_pair *pairs, size_t n_pairs, ocb_match *out, size_t out_capacity,
                               uint64_t *out_offsets, uint32_t *quality_order)
    {
        return match_pairs_tail(pairs, n_pairs, out, out_capacity, out_offsets, true, quality_order);
    }
    static int match_pairs_tail(const ocb_pair *pairs, size_t n_pairs, ocb_match *out, size_t out_capacity,
                                uint64_t *out_offsets, bool sorted, uint32_t *quality_order)
    {
        if (!out_offsets)
            return fail_invalid("null pointer");
        out_offsets[0] = 0;
        if (n_pairs == 0)
            return 0;
        if (!pairs || (!out && out_capacity))
            return fail_invalid("null pointer");
        if (n_pairs >= (1ull << 31))
            return fail_invalid("too many pairs");
        ThreadCtx &cx = t_ctx;
        int rc = cx.ensure();
        if (rc)
            return rc;
        cudaStream_t bulk = cx.bulk;
        std::vector<K1Problem> pr(n_pairs);
        memset(pr.data(), 0, sizeof(K1Problem) * n_pairs);
        std::vector<uint64_t> top_off(n_pairs + 1, 0);
        {
            std::lock_guard<std::mutex> lk(g_sets_mu);
            for (size_t p = 0; p < n_pairs; p++)
            {
                auto a = g_sets.find(pairs[p].query_set), b = g_sets.find(pairs[p].candidate_set);
                if (a == g_sets.end() || b == g_sets.end())
                {
                    set_last_error("unknown descriptor set in pair list");
                    return OCB_E_NOT_FOUND;
                }
                if (a->second.device != cx.device || b->second.device != cx.device)
                    return fail_invalid("descriptor set registered on another device");
                pr[p].q = static_cast<const uint4 *>(a->second.d_rows), pr[p].n_q = (uint32_t)a->second.n;
                pr[p].c = static_cast<const uint4 *>(b->second.d_rows), pr[p].n_c = (uint32_t)b->second.n;
                top_off[p + 1] = top_off[p] + a->second.n;
            }
        }
        const uint64_t rows_total = top_off[n_pairs];
        K1Plan plan = k1_plan(pr.data(), n_pairs, sm_count(cx.device));
        if ((uint64_t)plan.total_items >= 0x7FFFFFFFull)
            return fail_invalid("too many work items for one submission");
        const size_t table_bytes = sizeof(K1Problem) * n_pairs, k5_bytes = sizeof(K5Pair) * n_pairs;
        size_t state_total = 0;
        for (size_t p = 0; p < n_pairs; p++)
            state_total += align_up(k1_state_bytes(pr[p]), 16);
        // device block: [K1 table][K5 table][top-2 records of every query][K1 state][offsets][ticket][survivors]
        Carver cv;
        const size_t o_tab = cv.take(table_bytes), o_k5 = cv.take(k5_bytes), o_top = cv.take(rows_total * sizeof(ocb_top2));
        const size_t o_state = cv.take(state_total), o_off = cv.take((n_pairs + 1) * sizeof(uint64_t));
        const size_t o_ticket = cv.take(sizeof(uint32_t)), o_out = cv.take(rows_total * sizeof(ocb_match));
        // K7: sorted copy of the survivors, PROSAC order, scratch words for lists too long for shared memory
        const size_t o_sorted = cv.take(sorted ? rows_total * sizeof(ocb_match) : 0);
        const size_t o_qorder = cv.take(sorted && quality_order ? rows_total * sizeof(uint32_t) : 0);
        const size_t o_words = cv.take(sorted ? rows_total * sizeof(uint64_t) : 0);
        rc = cx.dev_reserve(cv.off);
        if (rc)
            return rc;
        const bool out_pinned = out_capacity && is_pinned_host(out), off_pinned = is_pinned_host(out_offsets);
        Carver sv;
        const size_t s_tab = sv.take(table_bytes), s_k5 = sv.take(k5_bytes);
        const size_t s_off = sv.take((n_pairs + 1) * sizeof(uint64_t));
        const size_t s_out = sv.take(out_pinned ? 0 : rows_total * sizeof(ocb_match));
        const bool qo_pinned = sorted && quality_order && is_pinned_host(quality_order);
        const size_t s_qo = sv.take(sorted && quality_order && !qo_pinned ? rows_total * sizeof(uint32_t) : 0);
        rc = cx.pinned_reserve(sv.off);
        if (rc)
            return rc;
        char *d = static_cast<char *>(cx.dev.p);
        char *hp = static_cast<char *>(cx.pinned.p);
        K5Pair *k5 = reinterpret_cast<K5Pair *>(hp + s_k5);
        size_t state_off = 0;
        for (size_t p = 0; p < n_pairs; p++)
        {
            pr[p].out = reinterpret_cast<ocb_top2 *>(d + o_top) + top_off[p];
            k1_bind_state(pr[p], d + o_state + state_off);
            state_off += align_up(k1_state_bytes(pr[p]), 16);
            k5[p].top = pr[p].out, k5[p].n_q = pr[p].n_q, k5[p].pad = 0;
        }
        if (state_total)
            OCB_CUDA(cudaMemsetAsync(d + o_state, 0, state_total, bulk));
        // both tables travel in one copy (they are adjacent in the staging image and on the device)
        if (o_k5 - o_tab != s_k5 - s_tab)
            return fail_invalid("internal: staging layout");
        memcpy(hp + s_tab, pr.data(), table_bytes);
        OCB_CUDA(cudaMemcpyAsync(d + o_tab, hp + s_tab, (s_k5 - s_tab) + k5_bytes, cudaMemcpyHostToDevice, bulk));
        static_assert(sizeof(K1Problem) % 8 == 0, "tables are carved at the same relative offsets on both sides");
        const K1Problem *d_tab = n_pairs > (size_t)K1_INLINE ? reinterpret_cast<const K1Problem *>(d + o_tab) : nullptr;
        rc = k1_launch(d_tab, pr.data(), n_pairs, plan, bulk);
        if (rc)
            return rc;
        rc = k5_ratio_compact(reinterpret_cast<const K5Pair *>(d + o_k5), n_pairs,
                              reinterpret_cast<unsigned long long *>(d + o_off),
                              reinterpret_cast<uint32_t *>(d + o_ticket), reinterpret_cast<ocb_match *>(d + o_out), bulk);
        if (rc)
            return rc;
        if (sorted)
        {
            uint32_t longest = 0;
            for (size_t p = 0; p < n_pairs; p++)
                longest = std::max(longest, pr[p].n_q);
            rc = k7_sort(reinterpret_cast<const unsigned long long *>(d + o_off), n_pairs, longest,
                         reinterpret_cast<const ocb_match *>(d + o_out), reinterpret_cast<ocb_match *>(d + o_sorted),
                         quality_order ? reinterpret_cast<uint32_t *>(d + o_qorder) : nullptr,
                         reinterpret_cast<unsigned long long *>(d + o_words), bulk);
            if (rc)
                return rc;
        }
        const size_t o_result = sorted ? o_sorted : o_out;
        // the offsets first (they say how many survivors there are), then exactly that many records
        OCB_CUDA(cudaMemcpyAsync(off_pinned ? (void *)out_offsets : (void *)(hp + s_off), d + o_off,
                                 (n_pairs + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, bulk));
        if ((rc = cx.wait_bulk()))
            return rc;
        if (!off_pinned)
            memcpy(out_offsets, hp + s_off, (n_pairs + 1) * sizeof(uint64_t));
        const uint64_t total = out_offsets[n_pairs];
        if (total > out_capacity)
            return fail_invalid("out_capacity is smaller than the number of surviving matches");
        if (total)
        {
            OCB_CUDA(cudaMemcpyAsync(out_pinned ? (void *)out : (void *)(hp + s_out), d + o_result,
                                     total * sizeof(ocb_match), cudaMemcpyDeviceToHost, bulk));
            if (sorted && quality_order)
                OCB_CUDA(cudaMemcpyAsync(qo_pinned ? (void *)quality_order : (void *)(hp + s_qo), d + o_qorder,
                                         total * sizeof(uint32_t), cudaMemcpyDeviceToHost, bulk));
            if ((rc = cx.wait_bulk()))
                return rc;
            if (!out_pinned)
                memcpy(out, hp + s_out, total * sizeof(ocb_match));
            if (sorted && quality_order && !qo_pinned)
                memcpy(quality_order, hp + s_qo, total * sizeof(uint32_t));
        }
        return 0;
    }

    // ---------------------------------------------------------------------------------------------------
    // K4: candidate lists
    // ---------------------------------------------------------------------------------------------------
    int ocb_match_lists_device(const void *d_q, const void *d_c, const void *d_list_query, const void *d_list_begin,
                               const void *d_list_candidates, size_t n_lists, void *d_out, void *stream)
    {
        if (n_lists >= 0xFFFFFFFFull)
            return fail_invalid("n_lists must fit in 32 bits");
        if (n_lists && (!d_q || !d_list_query || !d_list_begin || !d_out))
            return fail_invalid("null device pointer");
        if (!aligned16(d_q) || !aligned16(d_c) || (reinterpret_cast<uintptr_t>(d_out) & 7u) ||
            (reinterpret_cast<uintptr_t>(d_list_begin) & 7u))
            return fail_invalid("device rows must be 16-byte aligned (out / list_begin 8)");
        return k4_launch(d_q, d_c, static_cast<const uint32_t *>(d_list_query),
                         static_cast<const uint64_t *>(d_list_begin), static_cast<const uint32_t *>(d_list_candidates),
                         n_lists, static_cast<ocb_top2 *>(d_out), static_cast<cudaStream_t>(stream));
    }

    int ocb_match_lists(const uint64_t *q, size_t n_q_rows, const uint64_t *c, size_t n_c_rows,
                        const uint32_t *list_query, const uint64_t *list_begin, const uint32_t *list_candidates,
                        size_t n_lists, ocb_top2 *out)
    {
        if (n_q_rows >= 0xFFFFFFFFull || n_c_rows >= 0xFFFFFFFFull || n_lists >= 0xFFFFFFFFull)
            return fail_invalid("row / list counts must fit in 32 bits");
        if (n_lists == 0)
            return 0;
        if (!q || !list_query || !list_begin || !out)
            return fail_invalid("null pointer");
        // validate before anything reaches the device: a bad index would be a wild 64-byte read there
        const uint64_t total = list_begin[n_lists];
        if (list_begin[0] != 0)
            return fail_invalid("list_begin[0] must be 0");
        for (size_t l = 0; l < n_lists; l++)
        {
            if (list_begin[l + 1] < list_begin[l] || list_begin[l + 1] - list_begin[l] > OCB_MAX_LIST_LENGTH)
                return fail_invalid("list_begin must be non-decreasing, lists at most OCB_MAX_LIST_LENGTH long");
            if (list_query[l] >= n_q_rows)
                return fail_invalid("list_query index out of range");
        }
        if (total && (!c || !list_candidates))
            return fail_invalid("null pointer");
        {
            uint32_t worst = 0;
            for (uint64_t k = 0; k < total; k++)
                worst = std::max(worst, list_candidates[k]);
            if (total && worst >= n_c_rows)
                return fail_invalid("list_candidates index out of range");
        }
        ThreadCtx &cx = t_ctx;
        int rc = cx.ensure();
        if (rc)
            return rc;
        const size_t qb = n_q_rows * OCB_ROW_BYTES, cb = n_c_rows * OCB_ROW_BYTES;
        const size_t lqb = n_lists * sizeof(uint32_t), lbb = (n_lists + 1) * sizeof(uint64_t);
        const size_t lcb = (size_t)total * sizeof(uint32_t), ob = n_lists * sizeof(ocb_top2);
        Carver cv;
        const size_t o_q = cv.take(qb), o_c = cv.take(cb), o_lq = cv.take(lqb), o_lb = cv.take(lbb);
        const size_t o_lc = cv.take(lcb), o_out = cv.take(ob);
        rc = cx.dev_reserve(cv.off);
        if (rc)
            return rc;
        Carver sv;
        const size_t s_q = sv.take(qb), s_c = sv.take(cb), s_lq = sv.take(lqb), s_lb = sv.take(lbb);
        const size_t s_lc = sv.take(lcb), s_out = sv.take(ob);
        rc = cx.pinned_reserve(sv.off);
        if (rc)
            return rc;
        char *d = static_cast<char *>(cx.dev.p);
        if ((rc = upload(cx, d + o_q, q, qb, s_q)) || (rc = upload(cx, d + o_c, c, cb, s_c)) ||
            (rc = upload(cx, d + o_lq, list_query, lqb, s_lq)) || (rc = upload(cx, d + o_lb, list_begin, lbb, s_lb)) ||
            (rc = upload(cx, d + o_lc, list_candidates, lcb, s_lc)))
            return rc;
        rc = ocb_match_lists_device(d + o_q, d + o_c, d + o_lq, d + o_lb, d + o_lc, n_lists, d + o_out, cx.stream);
        if (rc)
            return rc;
        char *hp = static_cast<char *>(cx.pinned.p);
        const bool out_pinned = is_pinned_host(out);
        OCB_CUDA(cudaMemcpyAsync(out_pinned ? (void *)out : (void *)(hp + s_out), d + o_out, ob, cudaMemcpyDeviceToHost,
                                 cx.stream));
        OCB_CUDA(cudaStreamSynchronize(cx.stream));
        if (!out_pinned)
            memcpy(out, hp + s_out, ob);
        return 0;
    }

    // ---------------------------------------------------------------------------------------------------
    // K2 / K3
    // ---------------------------------------------------------------------------------------------------
    int ocb_prepare_correspondences_device(const void *d_corr7, const void *d_order, size_t n, void *d_corr4,
                                           void *d_pos, void *stream)
    {
        if (n >= 0xFFFFFFFFull)
            return fail_invalid("n must fit in 32 bits");
        if (n && (!d_corr7 || !d_corr4))
            return fail_invalid("null device pointer");
        if (reinterpret_cast<uintptr_t>(d_corr4) & 31u)
            return fail_invalid("d_corr4 must be 32-byte aligned");
        return k2_prepare(static_cast<const double *>(d_corr7), static_cast<const uint32_t *>(d_order), n,
                          static_cast<double *>(d_corr4), static_cast<uint32_t *>(d_pos),
                          static_cast<cudaStream_t>(stream));
    }

    int ocb_score_models_device(int kind, const void *d_models, size_t h, const void *d_corr4, const void *d_pos,
                                size_t n, double thr, void *d_score, void *d_count, void *d_inlier_bits, void *stream)
    {
        if (kind < 0 || kind > 2)
            return fail_invalid("kind");
        if (n >= 0xFFFFFFFFull || h >= 0xFFFFFFFFull)
            return fail_invalid("sizes must fit in 32 bits");
        if (h && (!d_models || !d_score || !d_count || (n && !d_corr4)))
            return fail_invalid("null device pointer");
        if (d_inlier_bits && d_pos)
            return fail_invalid("inlier bits with an evaluation order need scratch: use ocb_score_models");
        return k2_score(kind, static_cast<const double *>(d_models), h, static_cast<const double *>(d_corr4),
                        static_cast<const uint32_t *>(d_pos), n, thr, static_cast<double *>(d_score),
                        static_cast<uint32_t *>(d_count), static_cast<uint32_t *>(d_inlier_bits), nullptr,
                        static_cast<cudaStream_t>(stream));
    }

    int ocb_score_models(int kind, const double *models, size_t h, const double *corr, size_t n, double thr,
                         const uint32_t *order, double *score, uint32_t *count, uint32_t *inlier_bits)
    {
        if (kind < 0 || kind > 2)
            return fail_invalid("kind");
        if (n >= 0xFFFFFFFFull || h >= 0xFFFFFFFFull)
            return fail_invalid("sizes must fit in 32 bits");
        if (h == 0)
            return 0;
        if (!models || !score || !count || (n && !corr))
            return fail_invalid("null pointer");
        if (!order_in_range(order, n))
            return fail_invalid("order entry out of range");
        ThreadCtx &cx = t_ctx;
        int rc = cx.ensure();
        if (rc)
            return rc;
        const size_t words = (n + 31) / 32;
        const size_t mb = h * 18 * sizeof(double), cb = n * 7 * sizeof(double), ob = order ? n * sizeof(uint32_t) : 0;
        const size_t bb = inlier_bits ? h * words * sizeof(uint32_t) : 0;
        Carver cv;
        const size_t o_m = cv.take(mb), o_c7 = cv.take(cb), o_ord = cv.take(ob), o_c4 = cv.take(n * 32),
                     o_pos = cv.take(ob), o_sc = cv.take(h * sizeof(double)), o_cnt = cv.take(h * sizeof(uint32_t)),
                     o_bits = cv.take(bb), o_scr = cv.take(order ? bb : 0);
        rc = cx.dev_reserve(cv.off);
        if (rc)
            return rc;
        Carver sv;
        const size_t s_m = sv.take(mb), s_c7 = sv.take(cb), s_ord = sv.take(ob), s_sc = sv.take(h * sizeof(double)),
                     s_cnt = sv.take(h * sizeof(uint32_t)), s_bits = sv.take(bb);
        rc = cx.pinned_reserve(sv.off);
        if (rc)
            return rc;
        char *d = static_cast<char *>(cx.dev.p);
        char *hp = static_cast<char *>(cx.pinned.p);
        if ((rc = upload(cx, d + o_m, models, mb, s_m)))
            return rc;
        if ((rc = upload(cx, d + o_c7, corr, cb, s_c7)))
            return rc;
        if (order && (rc = upload(cx, d + o_ord, order, ob, s_ord)))
            return rc;
        uint32_t *d_pos = order ? reinterpret_cast<uint32_t *>(d + o_pos) : nullptr;
        rc = k2_prepare(reinterpret_cast<double *>(d + o_c7), order ? reinterpret_cast<uint32_t *>(d + o_ord) : nullptr,
                        n, reinterpret_cast<double *>(d + o_c4), d_pos, cx.stream);
        if (rc)
            return rc;
        rc = k2_score(kind, reinterpret_cast<double *>(d + o_m), h, reinterpret_cast<double *>(d + o_c4), d_pos, n, thr,
                      reinterpret_cast<double *>(d + o_sc), reinterpret_cast<uint32_t *>(d + o_cnt),
                      inlier_bits ? reinterpret_cast<uint32_t *>(d + o_bits) : nullptr,
                      reinterpret_cast<uint32_t *>(d + o_scr), cx.stream);
        if (rc)
            return rc;
        OCB_CUDA(cudaMemcpyAsync(hp + s_sc, d + o_sc, h * sizeof(double), cudaMemcpyDeviceToHost, cx.stream));
        OCB_CUDA(cudaMemcpyAsync(hp + s_cnt, d + o_cnt, h * sizeof(uint32_t), cudaMemcpyDeviceToHost, cx.stream));
        if (bb)
            OCB_CUDA(cudaMemcpyAsync(hp + s_bits, d + o_bits, bb, cudaMemcpyDeviceToHost, cx.stream));
        OCB_CUDA(cudaStreamSynchronize(cx.stream));
        memcpy(score, hp + s_sc, h * sizeof(double));
        memcpy(count, hp + s_cnt, h * sizeof(uint32_t));
        if (bb)
            memcpy(inlier_bits, hp + s_bits, bb);
        return 0;
    }

    // ---- correspondences resident for the calling thread ------------------------------------------------
    namespace
    {
    struct BoundLayout
    {
        size_t o_c7, o_nat, o_ord, o_pos, total;
    };
    BoundLayout bound_layout(size_t n)
    {
        Carver cv;
        BoundLayout L;
        L.o_c7 = cv.take(n * 7 * sizeof(double));
        L.o_nat = cv.take(n * 32);
        L.o_ord = cv.take(n * 32);
        L.o_pos = cv.take(n * sizeof(uint32_t));
        L.total = cv.off;
        return L;
    }
    } // namespace

    int ocb_corr_bind(const double *corr, size_t n, const uint32_t *order)
    {
        if (n >= 0xFFFFFFFFull)
            return fail_invalid("n must fit in 32 bits");
        if (n && !corr)
            return fail_invalid("corr");
        if (!order_in_range(order, n))
            return fail_invalid("order entry out of range");
        ThreadCtx &cx = t_ctx;
        int rc = cx.ensure();
        if (rc)
            return rc;
        cx.bound_valid = false;
        const BoundLayout L = bound_layout(n);
        if (L.total > cx.bound.cap)
        {
            OCB_CUDA(cudaStreamSynchronize(cx.stream));
            if (cx.bound.p)
                OCB_CUDA(cudaFree(cx.bound.p));
            cx.bound = Buf();
            const size_t cap = std::max(L.total * 2, (size_t)1 << 20);
            OCB_CUDA(cudaMalloc(&cx.bound.p, cap));
            cx.bound.cap = cap;
        }
        const size_t cb = n * 7 * sizeof(double), ob = order ? n * sizeof(uint32_t) : 0;
        Carver sv;
        const size_t s_c7 = sv.take(cb), s_ord = sv.take(ob);
        Carver dv;
        const size_t o_ord32 = dv.take(ob);
        if ((rc = cx.pinned_reserve(sv.off)) || (rc = cx.dev_reserve(dv.off)))
            return rc;
        char *b = static_cast<char *>(cx.bound.p);
        char *d = static_cast<char *>(cx.dev.p);
        if (n)
        {
            if ((rc = upload(cx, b + L.o_c7, corr, cb, s_c7)))
                return rc;
            if (order && (rc = upload(cx, d + o_ord32, order, ob, s_ord)))
                return rc;
            rc = k2_prepare(reinterpret_cast<double *>(b + L.o_c7), nullptr, n, reinterpret_cast<double *>(b + L.o_nat),
                            nullptr, cx.stream);
            if (!rc && order)
                rc = k2_prepare(reinterpret_cast<double *>(b + L.o_c7), reinterpret_cast<uint32_t *>(d + o_ord32), n,
                                reinterpret_cast<double *>(b + L.o_ord), reinterpret_cast<uint32_t *>(b + L.o_pos),
                                cx.stream);
            if (rc)
                return rc;
            // the order indices live in the per-call buffer: they must be consumed before the next call reuses it
            OCB_CUDA(cudaStreamSynchronize(cx.stream));
        }
        cx.bound_n = n;
        cx.bound_has_order = order != nullptr;
        cx.bound_valid = true;
        return 0;
    }

    int ocb_corr_unbind(void)
    {
        t_ctx.bound_valid = false;
        return 0;
    }

    int ocb_score_bound(int kind, const double *models, size_t h, double thr, int in_order, double *score,
                        uint32_t *count, uint32_t *inlier_bits)
    {
        if (kind < 0 || kind > 2)
            return fail_invalid("kind");
        if (h >= 0xFFFFFFFFull)
            return fail_invalid("sizes must fit in 32 bits");
        ThreadCtx &cx = t_ctx;
        int rc = cx.ensure();
        if (rc)
            return rc;
        if (!cx.bound_valid)
            return fail_invalid("no correspondences bound on this thread (ocb_corr_bind)");
        if (in_order && !cx.bound_has_order)
            return fail_invalid("correspondences were bound without an evaluation order");
        if (h == 0)
            return 0;
        if (!models || !score || !count)
            return fail_invalid("null pointer");
        const size_t n = cx.bound_n, words = (n + 31) / 32;
        const BoundLayout L = bound_layout(n);
        const size_t mb = h * 18 * sizeof(double), bb = inlier_bits ? h * words * sizeof(uint32_t) : 0;
        Carver cv;
        const size_t o_m = cv.take(mb), o_sc = cv.take(h * sizeof(double)), o_cnt = cv.take(h * sizeof(uint32_t)),
                     o_bits = cv.take(bb), o_scr = cv.take(in_order ? bb : 0);
        Carver sv;
        const size_t s_m = sv.take(mb), s_sc = sv.take(h * sizeof(double)), s_cnt = sv.take(h * sizeof(uint32_t)),
                     s_bits = sv.take(bb);
        if ((rc = cx.dev_reserve(cv.off)) || (rc = cx.pinned_reserve(sv.off)))
            return rc;
        char *b = static_cast<char *>(cx.bound.p);
        char *d = static_cast<char *>(cx.dev.p);
        char *hp = static_cast<char *>(cx.pinned.p);
        if ((rc = upload(cx, d + o_m, models, mb, s_m)))
            return rc;
        rc = k2_score(kind, reinterpret_cast<double *>(d + o_m), h,
                      reinterpret_cast<double *>(b + (in_order ? L.o_ord : L.o_nat)),
                      in_order ? reinterpret_cast<uint32_t *>(b + L.o_pos) : nullptr, n, thr,
                      reinterpret_cast<double *>(d + o_sc), reinterpret_cast<uint32_t *>(d + o_cnt),
                      inlier_bits ? reinterpret_cast<uint32_t *>(d + o_bits) : nullptr,
                      reinterpret_cast<uint32_t *>(d + o_scr), cx.stream);
        if (rc)
            return rc;
        OCB_CUDA(cudaMemcpyAsync(hp + s_sc, d + o_sc, h * sizeof(double), cudaMemcpyDeviceToHost, cx.stream));
        OCB_CUDA(cudaMemcpyAsync(hp + s_cnt, d + o_cnt, h * sizeof(uint32_t), cudaMemcpyDeviceToHost, cx.stream));
        if (bb)
            OCB_CUDA(cudaMemcpyAsync(hp + s_bits, d + o_bits, bb, cudaMemcpyDeviceToHost, cx.stream));
        OCB_CUDA(cudaStreamSynchronize(cx.stream));
        memcpy(score, hp + s_sc, h * sizeof(double));
        memcpy(count, hp + s_cnt, h * sizeof(uint32_t));
        if (bb)
            memcpy(inlier_bits, hp + s_bits, bb);
        return 0;
    }

    int ocb_fit_score_bound(const uint32_t *samples, size_t h, double thr, double *models_out, uint8_t *degenerate,
                            double *score, uint32_t *count)
    {
        if (h >= 0xFFFFFFFFull)
            return fail_invalid("sizes must fit in 32 bits");
        ThreadCtx &cx = t_ctx;
        int rc = cx.ensure();
        if (rc)
            return rc;
        if (!cx.bound_valid)
            return fail_invalid("no correspondences bound on this thread (ocb_corr_bind)");
        if (score && !cx.bound_has_order)
            return fail_invalid("correspondences were bound without an evaluation order");
        if (h == 0)
            return 0;
        if (!samples || !models_out || !degenerate || (score && !count))
            return fail_invalid("null pointer");
        const size_t n = cx.bound_n;
        for (size_t i = 0; i < h * 4; i++)
            if (samples[i] >= n)
                return fail_invalid("sample index out of range");
        const BoundLayout L = bound_layout(n);
        const size_t sb = h * 4 * sizeof(uint32_t), mb = h * 18 * sizeof(double);
        Carver cv; // device: [job][samples] in, [models][degenerate][score][count] out (one copy back)
        const size_t o_job = cv.take(sizeof(K3FitJob)), o_s = cv.take(sb);
        const size_t in_bytes = cv.off;
        const size_t o_m = cv.take(mb), o_dg = cv.take(h), o_sc = cv.take(score ? h * sizeof(double) : 0),
                     o_cnt = cv.take(score ? h * sizeof(uint32_t) : 0);
        const size_t out_bytes = cv.off - o_m;
        if ((rc = cx.dev_reserve(cv.off)) || (rc = cx.pinned_reserve(cv.off)))
            return rc;
        char *b = static_cast<char *>(cx.bound.p);
        char *d = static_cast<char *>(cx.dev.p);
        char *hp = static_cast<char *>(cx.pinned.p);
        K3FitJob job;
        job.c7 = reinterpret_cast<const double *>(b + L.o_c7);
        job.samples = reinterpret_cast<const uint32_t *>(d + o_s);
        job.models_out = reinterpret_cast<double *>(d + o_m);
        job.degenerate = reinterpret_cast<uint8_t *>(d + o_dg);
        job.h = (uint32_t)h, job.begin = 0;
        memcpy(hp + o_job, &job, sizeof job);
        memcpy(hp + o_s, samples, sb);
        OCB_CUDA(cudaMemcpyAsync(d, hp, in_bytes, cudaMemcpyHostToDevice, cx.stream));
        if ((rc = k3_fit_samples(reinterpret_cast<const K3FitJob *>(d + o_job), 1, (uint32_t)h, cx.stream)))
            return rc;
        if (score)
        {
            rc = k2_score(OCB_MODEL_HOMOGRAPHY, reinterpret_cast<double *>(d + o_m), h,
                          reinterpret_cast<double *>(b + L.o_ord), reinterpret_cast<uint32_t *>(b + L.o_pos), n, thr,
                          reinterpret_cast<double *>(d + o_sc), reinterpret_cast<uint32_t *>(d + o_cnt), nullptr, nullptr,
                          cx.stream);
            if (rc)
                return rc;
        }
        OCB_CUDA(cudaMemcpyAsync(hp + o_m, d + o_m, out_bytes, cudaMemcpyDeviceToHost, cx.stream));
        OCB_CUDA(cudaStreamSynchronize(cx.stream));
        memcpy(models_out, hp + o_m, mb);
        memcpy(degenerate, hp + o_dg, h);
        if (score)
        {
            memcpy(score, hp + o_sc, h * sizeof(double));
            memcpy(count, hp + o_cnt, h * sizeof(uint32_t));
        }
        return 0;
    }

    int ocb_refit_evaluate_bound(const uint32_t *refit_bits, double thr, double *model_out, double *score,
                                 uint32_t *count, uint32_t *inlier_bits)
    {
        ThreadCtx &cx = t_ctx;
        int rc = cx.ensure();
        if (rc)
            return rc;
        if (!cx.bound_valid)
            return fail_invalid("no correspondences bound on this thread (ocb_corr_bind)");
        if (!refit_bits || !model_out || !score || !count || !inlier_bits)
            return fail_invalid("null pointer");
        const size_t n = cx.bound_n, words = (n + 31) / 32;
        if (n == 0)
            return fail_invalid("empty correspondence set");
        const BoundLayout L = bound_layout(n);
        const size_t wb = words * sizeof(uint32_t);
        Carver cv; // device: [job][mask] in, [model][score][count][bits] out (one copy back), [system] scratch
        const size_t o_job = cv.take(sizeof(K3InlierJob)), o_mask = cv.take(wb);
        const size_t in_bytes = cv.off;
        const size_t o_m = cv.take(18 * sizeof(double)), o_sc = cv.take(sizeof(double)), o_cnt = cv.take(sizeof(uint32_t)),
                     o_bits = cv.take(wb);
        const size_t out_bytes = cv.off - o_m;
        const size_t io_bytes = cv.off;
        const size_t o_sys = cv.take(k3_inlier_scratch_bytes(n));
        if ((rc = cx.dev_reserve(cv.off)) || (rc = cx.pinned_reserve(io_bytes)))
            return rc;
        char *b = static_cast<char *>(cx.bound.p);
        char *d = static_cast<char *>(cx.dev.p);
        char *hp = static_cast<char *>(cx.pinned.p);
        K3InlierJob job;
        job.c7 = reinterpret_cast<const double *>(b + L.o_c7);
        job.bits = reinterpret_cast<const uint32_t *>(d + o_mask);
        job.P = reinterpret_cast<double *>(d + o_sys);
        job.model_out = reinterpret_cast<double *>(d + o_m);
        job.n = (uint32_t)n;
        memcpy(hp + o_job, &job, sizeof job);
        memcpy(hp + o_mask, refit_bits, wb);
        OCB_CUDA(cudaMemcpyAsync(d, hp, in_bytes, cudaMemcpyHostToDevice, cx.stream));
        if ((rc = k3_fit_inliers(reinterpret_cast<const K3InlierJob *>(d + o_job), 1, cx.stream)))
            return rc;
        // Model::evaluate: index order, inlier mask out
        rc = k2_score(OCB_MODEL_HOMOGRAPHY, reinterpret_cast<double *>(d + o_m), 1, reinterpret_cast<double *>(b + L.o_nat),
                      nullptr, n, thr, reinterpret_cast<double *>(d + o_sc), reinterpret_cast<uint32_t *>(d + o_cnt),
                      reinterpret_cast<uint32_t *>(d + o_bits), nullptr, cx.stream);
        if (rc)
            return rc;
        OCB_CUDA(cudaMemcpyAsync(hp + o_m, d + o_m, out_bytes, cudaMemcpyDeviceToHost, cx.stream));
        OCB_CUDA(cudaStreamSynchronize(cx.stream));
        memcpy(model_out, hp + o_m, 18 * sizeof(double));
        memcpy(score, hp + o_sc, sizeof(double));
        memcpy(count, hp + o_cnt, sizeof(uint32_t));
        memcpy(inlier_bits, hp + o_bits, wb);
        return 0;
    }

    int ocb_fit_homography(const double *corr, size_t n, const uint32_t *samples, size_t h, double *models_out,
                           uint8_t *degenerate)
    {
        if (h == 0)
            return 0;
        if (!corr || n == 0)
            return fail_invalid("corr");
        // fits need no evaluation order: bind in index order for the duration of the call
        int rc = ocb_corr_bind(corr, n, nullptr);
        if (rc)
            return rc;
        rc = ocb_fit_score_bound(samples, h, 0.0, models_out, degenerate, nullptr, nullptr);
        ocb_corr_unbind();
        return rc;
    }

    int ocb_residuals_bound(int kind, const double *model18, double *e)
    {
        if (kind < 0 || kind > 2)
            return fail_invalid("kind");
        ThreadCtx &cx = t_ctx;
        int rc = cx.ensure();
        if (rc)
            return rc;
        if (!cx.bound_valid)
            return fail_invalid("no correspondences bound on this thread (ocb_corr_bind)");
        const size_t n = cx.bound_n;
        if (n == 0)
            return 0;
        if (!model18 || !e)
            return fail_invalid("null pointer");
        const BoundLayout L = bound_layout(n);
        const size_t eb = n * sizeof(double);
        Carver cv;
        const size_t o_m = cv.take(18 * sizeof(double)), o_e = cv.take(eb);
        Carver sv;
        const size_t s_m = sv.take(18 * sizeof(double)), s_e = sv.take(eb);
        if ((rc = cx.dev_reserve(cv.off)) || (rc = cx.pinned_reserve(sv.off)))
            return rc;
        char *b = static_cast<char *>(cx.bound.p);
        char *d = static_cast<char *>(cx.dev.p);
        char *hp = static_cast<char *>(cx.pinned.p);
        if ((rc = upload(cx, d + o_m, model18, 18 * sizeof(double), s_m)))
            return rc;
        rc = k2_residuals(kind, reinterpret_cast<double *>(d + o_m), reinterpret_cast<double *>(b + L.o_c7), n,
                          reinterpret_cast<double *>(d + o_e), cx.stream);
        if (rc)
            return rc;
        OCB_CUDA(cudaMemcpyAsync(hp + s_e, d + o_e, eb, cudaMemcpyDeviceToHost, cx.stream));
        OCB_CUDA(cudaStreamSynchronize(cx.stream));
        memcpy(e, hp + s_e, eb);
        return 0;
    }

    // ---- many correspondence sets resident at once + request-table scoring --------------------------------
    int ocb_corr_bind_batch(const ocb_corr_set *sets, size_t count)
    {
        ThreadCtx &cx = t_ctx;
        int rc = cx.ensure();
        if (rc)
            return rc;
        cx.batch_valid = false;
        cx.batch_sets.clear();
        if (count && !sets)
            return fail_invalid("sets");
        Carver cv;
        cx.batch_sets.resize(count);
        for (size_t i = 0; i < count; i++)
        {
            if (sets[i].n >= 0xFFFFFFFFull || (sets[i].n && !sets[i].corr))
                return fail_invalid("correspondence set");
            if (!order_in_range(sets[i].order, sets[i].n))
                return fail_invalid("order entry out of range");
            cx.batch_sets[i].n = sets[i].n;
            cx.batch_sets[i].has_order = sets[i].order != nullptr && sets[i].n > 0;
            cx.batch_sets[i].o_c7 = cv.take(sets[i].n * 7 * sizeof(double));
            cx.batch_sets[i].o_ord = cv.take(cx.batch_sets[i].has_order ? sets[i].n * sizeof(uint32_t) : 0);
        }
        const size_t total = cv.off;
        if (total > cx.batch.cap)
        {
            OCB_CUDA(cudaStreamSynchronize(cx.stream));
            if (cx.batch.p)
                OCB_CUDA(cudaFree(cx.batch.p));
            cx.batch = Buf();
            const size_t cap = std::max(total + total / 2, (size_t)1 << 20);
            OCB_CUDA(cudaMalloc(&cx.batch.p, cap));
            cx.batch.cap = cap;
        }
        if (total)
        {
            // the staging image mirrors the device layout: one copy moves every set
            if ((rc = cx.pinned_reserve(total)))
                return rc;
            char *hp = static_cast<char *>(cx.pinned.p);
            for (size_t i = 0; i < count; i++)
            {
                if (sets[i].n == 0)
                    continue;
                memcpy(hp + cx.batch_sets[i].o_c7, sets[i].corr, sets[i].n * 7 * sizeof(double));
                if (cx.batch_sets[i].has_order)
                    memcpy(hp + cx.batch_sets[i].o_ord, sets[i].order, sets[i].n * sizeof(uint32_t));
            }
            OCB_CUDA(cudaMemcpyAsync(cx.batch.p, hp, total, cudaMemcpyHostToDevice, cx.stream));
            OCB_CUDA(cudaStreamSynchronize(cx.stream)); // the staging area is reused by the next call
        }
        cx.batch_valid = true;
        return 0;
    }

    int ocb_corr_bind_batch_matches(const ocb_match_set *sets, size_t count)
    {
        ThreadCtx &cx = t_ctx;
        int rc = cx.ensure();
        if (rc)
            return rc;
        cx.batch_valid = false;
        cx.batch_sets.clear();
        if (count && !sets)
            return fail_invalid("sets");
        // batch block (stays bound): per set [n][7] rows + evaluation order, laid out like ocb_corr_bind_batch;
        // per-call block: [K6 table][matches of every set]
        Carver cv, in_cv;
        cx.batch_sets.resize(count);
        std::vector<K6Set> tab(count);
        std::vector<size_t> o_match(count), o_order(count);
        const size_t o_tab = in_cv.take(count * sizeof(K6Set));
        uint32_t ctas = 0;
        {
            std::lock_guard<std::mutex> lk(g_sets_mu);
            for (size_t i = 0; i < count; i++)
            {
                const ocb_match_set &ms = sets[i];
                if (ms.n >= 0xFFFFFFFFull || (ms.n && !ms.matches))
                    return fail_invalid("match set");
                if (!order_in_range(ms.order, ms.n))
                    return fail_invalid("order entry out of range");
                K6Set &k = tab[i];
                memset(&k, 0, sizeof k);
                if (ms.n)
                {
                    auto a = g_sets.find(ms.set_1), b = g_sets.find(ms.set_2);
                    if (a == g_sets.end() || b == g_sets.end())
                    {
                        set_last_error("unknown image set in match set");
                        return OCB_E_NOT_FOUND;
                    }
                    if (a->second.device != cx.device || b->second.device != cx.device)
                        return fail_invalid("image set registered on another device");
                    if (!a->second.d_xy || !b->second.d_xy)
                        return fail_invalid("image set was registered without keypoints (ocb_register_images_batch)");
                    uint32_t worst_q = 0, worst_k = 0;
                    for (size_t m = 0; m < ms.n; m++)
                    {
                        worst_q = std::max(worst_q, ms.matches[m].query_k);
                        worst_k = std::max(worst_k, ms.matches[m].best_k);
                    }
                    if (worst_q >= a->second.n || worst_k >= b->second.n)
                        return fail_invalid("match position out of range");
                    k.xy1 = static_cast<const double2 *>(a->second.d_xy);
                    k.xy2 = static_cast<const double2 *>(b->second.d_xy);
                    k.cam1 = a->second.camera, k.cam2 = b->second.camera;
                }
                k.n = (uint32_t)ms.n;
                k.cta_begin = ctas;
                ctas += k6_set_ctas(k.n);
                cx.batch_sets[i].n = ms.n;
                cx.batch_sets[i].has_order = ms.order != nullptr && ms.n > 0;
                cx.batch_sets[i].o_c7 = cv.take(ms.n * 7 * sizeof(double));
                cx.batch_sets[i].o_ord = cv.take(cx.batch_sets[i].has_order ? ms.n * sizeof(uint32_t) : 0);
                o_match[i] = in_cv.take(ms.n * sizeof(ocb_match));
                o_order[i] = in_cv.take(cx.batch_sets[i].has_order ? ms.n * sizeof(uint32_t) : 0);
            }
        }
        const size_t total = cv.off;
        if (total > cx.batch.cap)
        {
            if ((rc = cx.wait_stream()))
                return rc;
            if (cx.batch.p)
                OCB_CUDA(cudaFree(cx.batch.p));
            cx.batch = Buf();
            const size_t cap = std::max(total + total / 2, (size_t)1 << 20);
            OCB_CUDA(cudaMalloc(&cx.batch.p, cap));
            cx.batch.cap = cap;
        }
        if (total)
        {
            // staging image: [per-call block][image of the batch block]. Matches and evaluation orders go up in the
            // per-call block (one copy); K6 writes the rows and moves the orders into the batch block, whose image
            // comes back in one copy when the caller wants the rows
            if ((rc = cx.dev_reserve(in_cv.off)) || (rc = cx.pinned_reserve(in_cv.off + total)))
                return rc;
            char *d = static_cast<char *>(cx.dev.p);
            char *b = static_cast<char *>(cx.batch.p);
            char *hp = static_cast<char *>(cx.pinned.p);
            char *hb = hp + in_cv.off;
            for (size_t i = 0; i < count; i++)
            {
                if (sets[i].n == 0)
                    continue;
                tab[i].matches = reinterpret_cast<const ocb_match *>(d + o_match[i]);
                tab[i].c7 = reinterpret_cast<double *>(b + cx.batch_sets[i].o_c7);
                memcpy(hp + o_match[i], sets[i].matches, sets[i].n * sizeof(ocb_match));
                if (cx.batch_sets[i].has_order)
                {
                    memcpy(hp + o_order[i], sets[i].order, sets[i].n * sizeof(uint32_t));
                    tab[i].order_src = reinterpret_cast<const uint32_t *>(d + o_order[i]);
                    tab[i].order_dst = reinterpret_cast<uint32_t *>(b + cx.batch_sets[i].o_ord);
                }
            }
            memcpy(hp + o_tab, tab.data(), count * sizeof(K6Set));
            OCB_CUDA(cudaMemcpyAsync(d, hp, in_cv.off, cudaMemcpyHostToDevice, cx.stream));
            if ((rc = k6_rays(reinterpret_cast<const K6Set *>(d + o_tab), count, ctas, cx.stream)))
                return rc;
            bool want_out = false;
            for (size_t i = 0; i < count; i++)
                want_out |= sets[i].n && sets[i].corr_out;
            if (want_out)
                OCB_CUDA(cudaMemcpyAsync(hb, b, total, cudaMemcpyDeviceToHost, cx.stream));
            if ((rc = cx.wait_stream()))
                return rc; // the staging area is reused by the next call
            if (want_out)
                for (size_t i = 0; i < count; i++)
                    if (sets[i].n && sets[i].corr_out)
                        memcpy(sets[i].corr_out, hb + cx.batch_sets[i].o_c7, sets[i].n * 7 * sizeof(double));
        }
        cx.batch_valid = true;
        return 0;
    }

    int ocb_image_to_3d(const double *xy, size_t n, const ocb_camera *camera, double *rays)
    {
        if (n >= 0xFFFFFFFFull)
            return fail_invalid("n must fit in 32 bits");
        if (n == 0)
            return 0;
        if (!xy || !camera || !rays)
            return fail_invalid("null pointer");
        ThreadCtx &cx = t_ctx;
        int rc = cx.ensure();
        if (rc)
            return rc;
        const size_t ib = n * 2 * sizeof(double), ob = n * 3 * sizeof(double);
        Carver cv;
        const size_t o_in = cv.take(ib), o_out = cv.take(ob);
        if ((rc = cx.dev_reserve(cv.off)) || (rc = cx.pinned_reserve(cv.off)))
            return rc;
        char *d = static_cast<char *>(cx.dev.p);
        char *hp = static_cast<char *>(cx.pinned.p);
        if ((rc = upload(cx, d + o_in, xy, ib, o_in)))
            return rc;
        if ((rc = k6_points(reinterpret_cast<const double *>(d + o_in), n, *camera, reinterpret_cast<double *>(d + o_out),
                            cx.stream)))
            return rc;
        OCB_CUDA(cudaMemcpyAsync(hp + o_out, d + o_out, ob, cudaMemcpyDeviceToHost, cx.stream));
        if ((rc = cx.wait_stream()))
                return rc;
        memcpy(rays, hp + o_out, ob);
        return 0;
    }

    int ocb_score_requests(const ocb_score_request *req, size_t count)
    {
        if (count == 0)
            return 0;
        if (!req)
            return fail_invalid("requests");
        ThreadCtx &cx = t_ctx;
        int rc = cx.ensure();
        if (rc)
            return rc;
        if (!cx.batch_valid)
            return fail_invalid("no correspondence batch bound on this thread (ocb_corr_bind_batch)");
        // layout of the per-call device block: [request table][fit jobs][refit jobs][models | samples | masks][outputs]
        // [refit scratch]; outputs (for fit requests: the fitted models and degeneracy flags too) are read back in one copy
        Carver in_cv, out_cv, scratch_cv;
        const size_t o_tab = in_cv.take(count * sizeof(K2Request));
        size_t n_fit = 0, n_refit = 0;
        for (size_t i = 0; i < count; i++)
        {
            n_fit += req[i].mode == OCB_REQ_FIT_SCORE_ORDERED;
            n_refit += req[i].mode == OCB_REQ_REFIT_EVALUATE;
        }
        const size_t o_jobs = in_cv.take(n_fit * sizeof(K3FitJob));
        const size_t o_rjobs = in_cv.take(n_refit * sizeof(K3InlierJob));
        std::vector<size_t> o_models(count), o_score(count), o_count(count), o_aux(count), o_flags(count), o_mask(count),
            o_scratch(count);
        for (size_t i = 0; i < count; i++)
        {
            const ocb_score_request &r = req[i];
            if (r.kind < 0 || r.kind > 2 || r.mode < 0 || r.mode > 4)
                return fail_invalid("request kind / mode");
            if (r.set >= cx.batch_sets.size())
                return fail_invalid("request references an unbound set");
            const ThreadCtx::BatchSet &bs = cx.batch_sets[r.set];
            const bool fit = r.mode == OCB_REQ_FIT_SCORE_ORDERED;
            const bool refit = r.mode == OCB_REQ_REFIT_EVALUATE;
            if (r.h == 0 || (!fit && !refit && !r.models) || (r.mode == 2 && (r.h != 1 || !r.residuals)) ||
                (r.mode != 2 && (!r.score || !r.count)) || ((r.mode == 1 || refit) && !r.inlier_bits))
                return fail_invalid("request pointers");
            if (refit)
            {
                if (r.kind != OCB_MODEL_HOMOGRAPHY || r.h != 1 || !r.refit_bits || !r.models_out)
                    return fail_invalid("refit request: homography only, h == 1; refit_bits and models_out are required");
                const size_t words = (bs.n + 31) / 32;
                o_mask[i] = in_cv.take(words * sizeof(uint32_t));
                o_models[i] = out_cv.take(18 * sizeof(double));
                o_score[i] = out_cv.take(sizeof(double));
                o_count[i] = out_cv.take(sizeof(uint32_t));
                o_aux[i] = out_cv.take(words * sizeof(uint32_t));
                o_scratch[i] = scratch_cv.take(k3_inlier_scratch_bytes(bs.n));
                continue;
            }
            if (fit && (r.kind != OCB_MODEL_HOMOGRAPHY || !r.samples || !r.models_out || !r.degenerate))
                return fail_invalid("fit request: homography only; samples, models_out and degenerate are required");
            if ((r.mode == 0 || fit) && !bs.has_order && bs.n)
                return fail_invalid("set was bound without an evaluation order");
            if (fit)
            {
                for (size_t k = 0; k < (size_t)r.h * 4; k++)
                    if (r.samples[k] >= bs.n)
                        return fail_invalid("sample index out of range");
                o_aux[i] = in_cv.take((size_t)r.h * 4 * sizeof(uint32_t)); // samples
                o_models[i] = out_cv.take((size_t)r.h * 18 * sizeof(double));
                o_flags[i] = out_cv.take(r.h);
                o_score[i] = out_cv.take((size_t)r.h * sizeof(double));
                o_count[i] = out_cv.take((size_t)r.h * sizeof(uint32_t));
                continue;
            }
            o_models[i] = in_cv.take((size_t)r.h * 18 * sizeof(double));
            const size_t words = (bs.n + 31) / 32;
            if (r.mode == 2)
                o_aux[i] = out_cv.take(bs.n * sizeof(double));
            else
            {
                o_score[i] = out_cv.take((size_t)r.h * sizeof(double));
                o_count[i] = out_cv.take((size_t)r.h * sizeof(uint32_t));
                o_aux[i] = out_cv.take(r.mode == 1 ? (size_t)r.h * words * sizeof(uint32_t) : 0);
            }
        }
        const size_t in_bytes = in_cv.off, out_bytes = out_cv.off;
        if ((rc = cx.dev_reserve(in_bytes + out_bytes + scratch_cv.off)) || (rc = cx.pinned_reserve(in_bytes + out_bytes)))
            return rc;
        char *d = static_cast<char *>(cx.dev.p);
        char *d_out = d + in_bytes;
        char *d_scratch = d_out + out_bytes;
        char *hp = static_cast<char *>(cx.pinned.p);
        char *b = static_cast<char *>(cx.batch.p);
        K2Request *tab = reinterpret_cast<K2Request *>(hp + o_tab);
        K3FitJob *jobs = reinterpret_cast<K3FitJob *>(hp + o_jobs);
        K3InlierJob *rjobs = reinterpret_cast<K3InlierJob *>(hp + o_rjobs);
        uint32_t ctas = 0, fit_total = 0;
        size_t j = 0, jr = 0;
        for (size_t i = 0; i < count; i++)
        {
            const ocb_score_request &r = req[i];
            const ThreadCtx::BatchSet &bs = cx.batch_sets[r.set];
            const bool fit = r.mode == OCB_REQ_FIT_SCORE_ORDERED;
            const bool refit = r.mode == OCB_REQ_REFIT_EVALUATE;
            K2Request q;
            memset(&q, 0, sizeof q);
            q.c7 = reinterpret_cast<const double *>(b + bs.o_c7);
            if (refit)
            {
                memcpy(hp + o_mask[i], r.refit_bits, (bs.n + 31) / 32 * sizeof(uint32_t));
                K3InlierJob &job = rjobs[jr++];
                job.c7 = q.c7;
                job.bits = reinterpret_cast<const uint32_t *>(d + o_mask[i]);
                job.P = reinterpret_cast<double *>(d_scratch + o_scratch[i]);
                job.model_out = reinterpret_cast<double *>(d_out + o_models[i]);
                job.n = (uint32_t)bs.n;
                q.models = job.model_out;
            }
            else if (fit)
            {
                memcpy(hp + o_aux[i], r.samples, (size_t)r.h * 4 * sizeof(uint32_t));
                K3FitJob &job = jobs[j++];
                job.c7 = q.c7;
                job.samples = reinterpret_cast<const uint32_t *>(d + o_aux[i]);
                job.models_out = reinterpret_cast<double *>(d_out + o_models[i]);
                job.degenerate = reinterpret_cast<uint8_t *>(d_out + o_flags[i]);
                job.h = r.h, job.begin = fit_total;
                fit_total += r.h;
                q.models = job.models_out;
            }
            else
            {
                memcpy(hp + o_models[i], r.models, (size_t)r.h * 18 * sizeof(double));
                q.models = reinterpret_cast<const double *>(d + o_models[i]);
            }
            q.order = ((r.mode == 0 || fit) && bs.has_order) ? reinterpret_cast<const uint32_t *>(b + bs.o_ord) : nullptr;
            q.thr = r.thr;
            q.h = r.h, q.n = (uint32_t)bs.n, q.words = (uint32_t)((bs.n + 31) / 32);
            // a fit request is scored like OCB_REQ_SCORE_ORDERED, a refit request evaluated like OCB_REQ_EVALUATE
            q.kind = r.kind, q.mode = fit ? 0 : (refit ? 1 : r.mode);
            if (r.mode == 2)
                q.e = reinterpret_cast<double *>(d_out + o_aux[i]);
            else
            {
                q.score = reinterpret_cast<double *>(d_out + o_score[i]);
                q.count = reinterpret_cast<uint32_t *>(d_out + o_count[i]);
                q.bits = (r.mode == 1 || refit) ? reinterpret_cast<uint32_t *>(d_out + o_aux[i]) : nullptr;
            }
            q.cta_begin = ctas;
            ctas += k2_request_ctas(q);
            tab[i] = q;
        }
        OCB_CUDA(cudaMemcpyAsync(d, hp, in_bytes, cudaMemcpyHostToDevice, cx.stream));
        if (n_fit && (rc = k3_fit_samples(reinterpret_cast<const K3FitJob *>(d + o_jobs), n_fit, fit_total, cx.stream)))
            return rc;
        if (n_refit && (rc = k3_fit_inliers(reinterpret_cast<const K3InlierJob *>(d + o_rjobs), n_refit, cx.stream)))
            return rc;
        if ((rc = k2_run_requests(reinterpret_cast<const K2Request *>(d + o_tab), count, ctas, cx.stream)))
            return rc;
        char *hout = hp + in_bytes;
        if (out_bytes)
            OCB_CUDA(cudaMemcpyAsync(hout, d_out, out_bytes, cudaMemcpyDeviceToHost, cx.stream));
        if ((rc = cx.wait_stream()))
                return rc;
        for (size_t i = 0; i < count; i++)
        {
            const ocb_score_request &r = req[i];
            const size_t n = cx.batch_sets[r.set].n, words = (n + 31) / 32;
            if (r.mode == 2)
                memcpy(r.residuals, hout + o_aux[i], n * sizeof(double));
            else
            {
                memcpy(r.score, hout + o_score[i], (size_t)r.h * sizeof(double));
                memcpy(r.count, hout + o_count[i], (size_t)r.h * sizeof(uint32_t));
                if (r.mode == 1 || r.mode == OCB_REQ_REFIT_EVALUATE)
                    memcpy(r.inlier_bits, hout + o_aux[i], (size_t)r.h * words * sizeof(uint32_t));
                if (r.mode == OCB_REQ_REFIT_EVALUATE)
                    memcpy(r.models_out, hout + o_models[i], 18 * sizeof(double));
                if (r.mode == OCB_REQ_FIT_SCORE_ORDERED)
                {
                    memcpy(r.models_out, hout + o_models[i], (size_t)r.h * 18 * sizeof(double));
                    memcpy(r.degenerate, hout + o_flags[i], r.h);
                }
            }
        }
        return 0;
    }

    int ocb_residuals(int kind, const double *model18, const double *corr, size_t n, double *e)
    {
        if (kind < 0 || kind > 2)
            return fail_invalid("kind");
        if (n >= 0xFFFFFFFFull)
            return fail_invalid("n must fit in 32 bits");
        if (n == 0)
            return 0;
        if (!model18 || !corr || !e)
            return fail_invalid("null pointer");
        ThreadCtx &cx = t_ctx;
        int rc = cx.ensure();
        if (rc)
            return rc;
        const size_t cb = n * 7 * sizeof(double), eb = n * sizeof(double);
        Carver cv;
        const size_t o_m = cv.take(18 * sizeof(double)), o_c7 = cv.take(cb), o_e = cv.take(eb);
        rc = cx.dev_reserve(cv.off);
        if (rc)
            return rc;
        Carver sv;
        const size_t s_m = sv.take(18 * sizeof(double)), s_c7 = sv.take(cb), s_e = sv.take(eb);
        rc = cx.pinned_reserve(sv.off);
        if (rc)
            return rc;
        char *d = static_cast<char *>(cx.dev.p);
        char *hp = static_cast<char *>(cx.pinned.p);
        if ((rc = upload(cx, d + o_m, model18, 18 * sizeof(double), s_m)))
            return rc;
        if ((rc = upload(cx, d + o_c7, corr, cb, s_c7)))
            return rc;
        rc = k2_residuals(kind, reinterpret_cast<double *>(d + o_m), reinterpret_cast<double *>(d + o_c7), n,
                          reinterpret_cast<double *>(d + o_e), cx.stream);
        if (rc)
            return rc;
        OCB_CUDA(cudaMemcpyAsync(hp + s_e, d + o_e, eb, cudaMemcpyDeviceToHost, cx.stream));
        OCB_CUDA(cudaStreamSynchronize(cx.stream));
        memcpy(e, hp + s_e, eb);
        return 0;
    }

} // extern "C"
