// Internal helpers shared by the kernels and the C-ABI layer of libocb.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <cstdio>
#include <string>

#include "../../include/ocb.h"

namespace ocb
{

// ---- error plumbing ---------------------------------------------------------------------------------
void set_last_error(const std::string &msg);
int fail_cuda(cudaError_t e, const char *what, const char *file, int line);
int fail_invalid(const char *what);

#define OCB_CUDA(call)                                                                                                 \
    do                                                                                                                 \
    {                                                                                                                  \
        cudaError_t e__ = (call);                                                                                      \
        if (e__ != cudaSuccess)                                                                                        \
            return ::ocb::fail_cuda(e__, #call, __FILE__, __LINE__);                                                   \
    } while (0)

extern std::atomic<uint64_t> g_kernel_launches;
inline void count_launch(uint64_t n = 1)
{
    g_kernel_launches.fetch_add(n, std::memory_order_relaxed);
}

// ---- tuning options -----------------------------------------------------------------------------------
struct Options
{
    int k1_variant = 0;       // 0 = default (see hamming_top2.cu variant table)
    int k1_items_per_sm = 32; // target work items per SM when splitting the candidate axis
    int k2_variant = 0;       // k2_score form: 0 = by size, 1 = one hypothesis group per CTA, 2 = four in lock-step
    int k2_hg = 0;            // hypotheses per CTA of k2_score; 0 = balance the SMs (k2_pick_group)
    int k1_update = 0;      // 0 = choose by candidate-run length, 1 = vote-and-skip, 2 = branch-free
    int k1_bf_rows = 1 << 20; // runs shorter than this use the branch-free update (measured: it wins at every length)
    int k1_engine = 0;        // single-pair search: 0 = by size, 1 = integer pipes (K1), 2 = tensor cores (K1T)
    int k1t_variant = 0;      // K1T search kernel: 0 = default, 1 = first form, 2 / 3 = second form, one / two query tiles per CTA
};
Options &options();

int sm_count(int device);

// ---- K1 launch interface (hamming_top2.cu) ----------------------------------------------------------------
// One (query set, candidate set) problem of a launch. All pointers are device pointers.
struct K1Problem
{
    const uint4 *q;             // [n_q][4] uint4 = 64-byte rows
    const uint4 *c;             // [n_c][4]
    ocb_top2 *out;              // [n_q]
    // state, zero before the launch (see k1_state_bytes / k1_bind_state):
    uint32_t *counters;         // [q_tiles + splits] tickets; used when splits > 1 or col_out
    unsigned long long *best64; // [n_q] complemented (distance, position) of the best so far; splits > 1
    uint32_t *sec32;            // [n_q] complemented second-best distance; splits > 1
    unsigned long long *col64;  // [n_c] complemented (distance, query) keys; cross-check only
    uint32_t *col_out;          // [n_c] best query position per candidate (cross-check), or nullptr
    uint32_t n_q, n_c;          //
    uint32_t q_tiles;           // ceil(n_q / queries per CTA)
    uint32_t splits;            // candidate-axis splits
    uint32_t rows_per_split;    // candidate rows per split
    uint32_t item_begin;        // index of this problem's first work item in the launch
};

constexpr int K1_INLINE = 2; // problems that fit in the kernel parameters
struct K1Inline
{
    K1Problem p[K1_INLINE];
};

struct K1Plan
{
    uint32_t total_items = 0;
    bool any_col = false;
};
// Fills q_tiles / splits / rows_per_split / item_begin of each problem (n_q, n_c, col_out must be set).
K1Plan k1_plan(K1Problem *problems, size_t n, int sms, int items_per_sm = 0); // 0: options().k1_items_per_sm
size_t k1_state_bytes(const K1Problem &P);          // bytes of zero-initialised state a planned problem needs
void k1_bind_state(K1Problem &P, void *d_state);    // point counters / best64 / sec32 / col64 into that block
// Enqueues the single K1 kernel for `n` problems. d_problems = device copy of the table (ignored when
// n <= K1_INLINE: the table then travels in the kernel parameters).
int k1_launch(const K1Problem *d_problems, const K1Problem *h_problems, size_t n, const K1Plan &plan,
              cudaStream_t stream);
int k1_queries_per_cta();

// ---- K1T launch interface (hamming_tensor.cu): the same search on the tensor cores, one pair per call ------------
bool k1t_supports(size_t n1, size_t n2);
size_t k1t_workspace_bytes(size_t n1, size_t n2, bool col);
int k1t_launch(const void *d_q, size_t n1, const void *d_c, size_t n2, ocb_top2 *d_out, uint32_t *d_col_best_q,
               void *d_workspace, int sms, cudaStream_t stream);

// ---- K5 / K6 launch interface (link_tail.cu) ----------------------------------------------------------------
struct K5Pair
{
    const ocb_top2 *top; // [n_q] records of one pair (device)
    uint32_t n_q, pad;
};
// counts the ratio-test survivors per pair, turns the counts into exclusive offsets (d_offsets [n_pairs + 1]) and writes
// the survivors of pair p in query order at d_out + d_offsets[p]
int k5_ratio_compact(const K5Pair *d_pairs, size_t n_pairs, unsigned long long *d_offsets, uint32_t *d_ticket,
                     ocb_match *d_out, cudaStream_t stream);
// K7: per pair, d_out[offsets[p] ..) = d_in[offsets[p] ..) in the order of the reference's std::sort by distance
// (descending); d_quality_order (nullable) = the positions of those sorted matches in the order of the reference's
// std::sort by quality (ascending). d_scratch: one 64-bit word per record, used by pairs too long for shared memory.
int k7_sort(const unsigned long long *d_offsets, size_t n_pairs, uint32_t max_rows_per_pair, const ocb_match *d_in,
            ocb_match *d_out, uint32_t *d_quality_order, unsigned long long *d_scratch, cudaStream_t stream);
struct K6Set
{
    const double2 *xy1, *xy2; // keypoint locations of the two registered sets
    const ocb_match *matches; // [n] sorted matches (device)
    double *c7;               // [n][7] out
    const uint32_t *order_src; // nullable: evaluation order as uploaded ...
    uint32_t *order_dst;       // ... moved next to the rows (the layout ocb_score_requests reads)
    ocb_camera cam1, cam2;
    uint32_t n, cta_begin;
};
uint32_t k6_set_ctas(uint32_t n);
int k6_rays(const K6Set *d_sets, size_t n_sets, uint32_t total_ctas, cudaStream_t stream);
int k6_points(const double *d_xy, size_t n, const ocb_camera &cam, double *d_rays, cudaStream_t stream);

// ---- K4 launch interface (hamming_lists.cu) ----------------------------------------------------------------
int k4_launch(const void *d_q_rows, const void *d_c_rows, const uint32_t *d_list_query, const uint64_t *d_list_begin,
              const uint32_t *d_list_candidates, size_t n_lists, ocb_top2 *d_out, cudaStream_t stream);

// ---- K2/K3 launch interface (score_models.cu) ---------------------------------------------------------------
int k2_prepare(const double *d_corr7, const uint32_t *d_order, size_t n, double *d_corr4, uint32_t *d_pos,
               cudaStream_t stream);
int k2_score(int kind, const double *d_models, size_t h, const double *d_corr4, const uint32_t *d_pos, size_t n,
             double thr, double *d_score, uint32_t *d_count, uint32_t *d_bits, uint32_t *d_bits_scratch,
             cudaStream_t stream);
int k2_residuals(int kind, const double *d_model18, const double *d_corr7, size_t n, double *d_e,
                 cudaStream_t stream);

// One entry of the request table of k2_requests_kernel (all pointers are device pointers).
struct K2Request
{
    const double *models;  // [h][18] (mode 2: one model)
    const double *c7;      // [n][7] correspondences as given
    const uint32_t *order; // mode 0: evaluation order; nullptr = index order
    double *score;         // [h]
    uint32_t *count;       // [h]
    uint32_t *bits;        // mode 1: [h][words] inlier masks in index order; else nullptr
    double *e;             // mode 2: [n] residuals
    double thr;
    uint32_t h, n, words, cta_begin;
    int32_t kind, mode; // mode 0 = score in evaluation order, 1 = evaluate (index order + bits), 2 = residuals
};
uint32_t k2_request_ctas(const K2Request &rq);
uint32_t k2_pick_group(size_t h, int sms, int resident);
int k2_run_requests(const K2Request *d_requests, size_t n_requests, uint32_t total_ctas, cudaStream_t stream);

// ---- K3 launch interface (fit_models.cu) ---------------------------------------------------------------------
// One batch of minimal-sample homography fits against one correspondence set (all pointers are device pointers).
struct K3FitJob
{
    const double *c7;        // [n][7] correspondences as given
    const uint32_t *samples; // [h][4] correspondence indices
    double *models_out;      // [h][18]
    uint8_t *degenerate;     // [h] or nullptr
    uint32_t h, begin;       // begin = index of this job's first hypothesis in the launch
};
int k3_fit_samples(const K3FitJob *d_jobs, size_t n_jobs, uint32_t total, cudaStream_t stream);
// One all-inlier refit (homography_model::fitInliers): the correspondences whose bit is set in `bits` (index order).
struct K3InlierJob
{
    const double *c7;     // [n][7] correspondences as given
    const uint32_t *bits; // [ceil(n/32)] inlier mask in index order
    double *P;            // scratch for the (2m+1) x 9 system: room for (2n+1) * 9 doubles
    double *model_out;    // [18]
    uint32_t n;
};
size_t k3_inlier_scratch_bytes(size_t n);
int k3_fit_inliers(const K3InlierJob *d_jobs, size_t n_jobs, cudaStream_t stream);

// ---- PTX helpers: mbarrier + 1-D bulk (TMA) copies -----------------------------------------------------------
#if defined(__CUDACC__)
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n"
                 ".reg .pred p;\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
                 "selp.u32 %0, 1, 0, p;\n"
                 "}\n"
                 : "=r"(ok)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity))
    {
    }
}
// cp.async.bulk global -> shared::cta, completion signalled on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// orders this thread's generic-proxy shared-memory writes before later async-proxy (bulk copy) accesses
__device__ __forceinline__ void fence_proxy_async_smem()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// atomics on addresses known to be global (a generic-address atomicMax compiles to an address-space test plus a
// shared-memory CAS loop next to the global ATOM)
__device__ __forceinline__ void red_max_u64_global(unsigned long long *p, unsigned long long v)
{
    asm volatile("red.global.max.u64 [%0], %1;" ::"l"(__cvta_generic_to_global(p)), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long atom_max_u64_global(unsigned long long *p, unsigned long long v)
{
    unsigned long long old;
    asm volatile("atom.global.max.u64 %0, [%1], %2;" : "=l"(old) : "l"(__cvta_generic_to_global(p)), "l"(v) : "memory");
    return old;
}
__device__ __forceinline__ void red_max_u32_global(uint32_t *p, uint32_t v)
{
    asm volatile("red.global.max.u32 [%0], %1;" ::"l"(__cvta_generic_to_global(p)), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t lop3_xor3(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t lop3_maj(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
#endif

} // namespace ocb
