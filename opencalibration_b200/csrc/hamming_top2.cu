// K1 -- brute-force Hamming top-2 over 512-bit descriptor rows (sm_100a).
//
// Replaces the loop nest of match_features_subset (reference src/match/match_features.cpp:71-93):
// for every query row the first candidate position at minimum Hamming distance, that distance, and the
// second-smallest distance counted with multiplicity (tie rule of :80-92).
//
// Mapping to the hardware
//   * one work item = one CTA = (query tile of 128*Q rows) x (a contiguous range of candidate rows);
//     every thread keeps Q query rows in registers (16 x u32 each, loaded as 128-bit vectors);
//   * candidate rows stream through shared memory in TILE_C-row tiles, copied by 1-D bulk TMA
//     (cp.async.bulk + mbarrier, SASS UBLKCP) into a K1_STAGES-deep ring, and are read back as warp-uniform
//     LDS.128 broadcasts, so one shared-memory read feeds 32 lanes x Q comparisons;
//   * distance = XOR + POPC on the integer pipes. POPC issues on the quarter-rate XU pipe, LOP3 on the ALU pipe
//     and IMAD on the FMA pipe, so part of the 16 popcounts per comparison is traded for carry-save adders
//     (F full adders = 2 LOP3 each, each removes one POPC) to balance the three pipes; the weighted sum of the
//     remaining popcounts is accumulated with IMADs straight into a packed key
//           key = (distance << 20) | candidate_index_within_item
//     whose unsigned order is exactly the reference's (distance, position) order;
//   * per query the two smallest keys are kept: m1 gives best distance + first position, m2's distance is the
//     reference's second_best (a later equal distance has a larger key, i.e. multiplicity is preserved).
//     Updates are rare after the first few candidates, so the common path is one compare per comparison and a
//     warp-uniform branch around the 3-op min/max update;
//   * when a single pair cannot fill 148 SMs the candidate axis is split across CTAs; every split writes its
//     (m1,m2) per query and a small merge kernel takes the top-2 of the union in position order.
#include "ocb_internal.cuh"

#include <algorithm>
#include <cstring>
#include <vector>

namespace ocb
{

constexpr int K1_THREADS = 128;
constexpr int K1_TILE_C = 64; // candidate rows per shared-memory tile (4 KB)
constexpr int K1_STAGES = 4;
constexpr int K1_KEY_SHIFT = 20; // candidates per work item < 2^20
constexpr uint32_t K1_KEY_IDX_MASK = (1u << K1_KEY_SHIFT) - 1;
constexpr uint32_t K1_MAX_TILES_PER_SPLIT = (1u << K1_KEY_SHIFT) / K1_TILE_C;

// ----------------------------------------------------------------------------------------------------------
// distance key: key0 + (popcount(q ^ c) << 20), with F carry-save full adders in front of the popcounts
// ----------------------------------------------------------------------------------------------------------
template <int F> __device__ __forceinline__ uint32_t hamming_key(const uint32_t (&q)[16], const uint4 (&c)[4], uint32_t key0)
{
    constexpr int F1 = F < 7 ? F : 7;                      // adders on weight-1 words (16 -> 16-2*F1)
    constexpr int F2 = F <= 7 ? 0 : (F - 7 < 3 ? F - 7 : 3); // adders on weight-2 words
    constexpr int F4 = F <= 10 ? 0 : 1;                    // adder on weight-4 words
    uint32_t w1[16 + 7];
    uint32_t w2[7 + 3 + 1];
    uint32_t w4[3 + 1 + 1];
    uint32_t w8[1 + 1];
    w1[0] = q[0] ^ c[0].x, w1[1] = q[1] ^ c[0].y, w1[2] = q[2] ^ c[0].z, w1[3] = q[3] ^ c[0].w;
    w1[4] = q[4] ^ c[1].x, w1[5] = q[5] ^ c[1].y, w1[6] = q[6] ^ c[1].z, w1[7] = q[7] ^ c[1].w;
    w1[8] = q[8] ^ c[2].x, w1[9] = q[9] ^ c[2].y, w1[10] = q[10] ^ c[2].z, w1[11] = q[11] ^ c[2].w;
    w1[12] = q[12] ^ c[3].x, w1[13] = q[13] ^ c[3].y, w1[14] = q[14] ^ c[3].z, w1[15] = q[15] ^ c[3].w;
    int h1 = 0, t1 = 16, h2 = 0, t2 = 0, h4 = 0, t4 = 0, t8 = 0;
#pragma unroll
    for (int f = 0; f < F1; f++)
    {
        w1[t1++] = lop3_xor3(w1[h1], w1[h1 + 1], w1[h1 + 2]);
        w2[t2++] = lop3_maj(w1[h1], w1[h1 + 1], w1[h1 + 2]);
        h1 += 3;
    }
#pragma unroll
    for (int f = 0; f < F2; f++)
    {
        w2[t2++] = lop3_xor3(w2[h2], w2[h2 + 1], w2[h2 + 2]);
        w4[t4++] = lop3_maj(w2[h2], w2[h2 + 1], w2[h2 + 2]);
        h2 += 3;
    }
#pragma unroll
    for (int f = 0; f < F4; f++)
    {
        w4[t4++] = lop3_xor3(w4[h4], w4[h4 + 1], w4[h4 + 2]);
        w8[t8++] = lop3_maj(w4[h4], w4[h4 + 1], w4[h4 + 2]);
        h4 += 3;
    }
    uint32_t key = key0;
#pragma unroll
    for (int i = h1; i < t1; i++)
        key += (uint32_t)__popc(w1[i]) * (1u << K1_KEY_SHIFT);
#pragma unroll
    for (int i = h2; i < t2; i++)
        key += (uint32_t)__popc(w2[i]) * (2u << K1_KEY_SHIFT);
#pragma unroll
    for (int i = h4; i < t4; i++)
        key += (uint32_t)__popc(w4[i]) * (4u << K1_KEY_SHIFT);
#pragma unroll
    for (int i = 0; i < t8; i++)
        key += (uint32_t)__popc(w8[i]) * (8u << K1_KEY_SHIFT);
    return key;
}

__device__ __forceinline__ const K1Problem *find_problem(const K1Problem *__restrict__ problems, uint32_t n, uint32_t item)
{
    uint32_t lo = 0, hi = n - 1;
    while (lo < hi)
    {
        const uint32_t mid = (lo + hi + 1) >> 1;
        if (problems[mid].item_begin <= item)
            lo = mid;
        else
            hi = mid - 1;
    }
    return problems + lo;
}

__device__ __forceinline__ ocb_top2 finish_keys(uint32_t m1, uint32_t m2, uint32_t c_begin)
{
    ocb_top2 r;
    if (m1 == 0xFFFFFFFFu)
    {
        r.best_k = 0; // feature_match best_match{i, 0, inf} (match_features.cpp:74)
        r.best_d = OCB_DIST_INF;
    }
    else
    {
        r.best_k = c_begin + (m1 & K1_KEY_IDX_MASK);
        r.best_d = (uint16_t)(m1 >> K1_KEY_SHIFT);
    }
    r.second_d = m2 == 0xFFFFFFFFu ? (uint16_t)OCB_DIST_INF : (uint16_t)(m2 >> K1_KEY_SHIFT);
    return r;
}

template <int Q, int F, int MINB>
__global__ void __launch_bounds__(K1_THREADS, MINB)
    k1_top2_kernel(const __grid_constant__ K1Inline inl, const K1Problem *__restrict__ problems, uint32_t n_problems)
{
    __shared__ alignas(128) uint4 tile[K1_STAGES][K1_TILE_C * 4];
    __shared__ alignas(8) uint64_t full_bar[K1_STAGES];

    const uint32_t tid = threadIdx.x;
    // <= K1_INLINE problems travel in the kernel parameters (no table upload on the single-pair path)
    const K1Problem *pp = problems ? find_problem(problems, n_problems, blockIdx.x)
                                   : &inl.p[(n_problems > 1 && blockIdx.x >= inl.p[1].item_begin) ? 1 : 0];
    const uint32_t n_q = pp->n_q, n_c = pp->n_c, q_tiles = pp->q_tiles;
    const uint32_t local = blockIdx.x - pp->item_begin;
    const uint32_t split = local / q_tiles;
    const uint32_t qtile = local - split * q_tiles;
    const uint32_t tiles_per_split = pp->tiles_per_split;
    const uint32_t c_begin = split * tiles_per_split * K1_TILE_C;
    const uint32_t c_end = min(n_c, c_begin + tiles_per_split * K1_TILE_C);
    const uint32_t c_cnt = c_end > c_begin ? c_end - c_begin : 0;
    const uint32_t ntiles = (c_cnt + K1_TILE_C - 1) / K1_TILE_C;
    const uint4 *__restrict__ cand = pp->c + (size_t)c_begin * 4;

    if (tid == 0)
    {
#pragma unroll
        for (int s = 0; s < K1_STAGES; s++)
            mbar_init(&full_bar[s], 1);
        mbar_fence_init();
    }
    __syncthreads();

    auto issue = [&](uint32_t t) {
        const uint32_t s = t % K1_STAGES;
        const uint32_t rows = min((uint32_t)K1_TILE_C, c_cnt - t * K1_TILE_C);
        const uint32_t bytes = rows * OCB_ROW_BYTES;
        mbar_expect_tx(&full_bar[s], bytes);
        bulk_g2s(&tile[s][0], cand + (size_t)t * K1_TILE_C * 4, bytes, &full_bar[s]);
    };
    if (tid == 0)
    {
        for (uint32_t t = 0; t < (uint32_t)K1_STAGES && t < ntiles; t++)
            issue(t);
    }

    // query rows -> registers (coalesced 128-bit loads; rows past n_q read row n_q-1 and are never written back)
    uint32_t q[Q][16];
    const uint32_t q_base = qtile * (K1_THREADS * Q);
#pragma unroll
    for (int j = 0; j < Q; j++)
    {
        uint32_t qi = q_base + j * K1_THREADS + tid;
        qi = qi < n_q ? qi : n_q - 1;
        const uint4 *row = pp->q + (size_t)qi * 4;
#pragma unroll
        for (int v = 0; v < 4; v++)
        {
            const uint4 x = __ldg(row + v);
            q[j][4 * v + 0] = x.x, q[j][4 * v + 1] = x.y, q[j][4 * v + 2] = x.z, q[j][4 * v + 3] = x.w;
        }
    }
    uint32_t m1[Q], m2[Q];
#pragma unroll
    for (int j = 0; j < Q; j++)
        m1[j] = m2[j] = 0xFFFFFFFFu;

    for (uint32_t t = 0; t < ntiles; t++)
    {
        const uint32_t s = t % K1_STAGES;
        mbar_wait(&full_bar[s], (t / K1_STAGES) & 1);
        const uint32_t rows = min((uint32_t)K1_TILE_C, c_cnt - t * K1_TILE_C);
        const uint4 *__restrict__ tl = &tile[s][0];
        const uint32_t k0 = t * K1_TILE_C;
#pragma unroll 2
        for (uint32_t cc = 0; cc < rows; cc++)
        {
            uint4 c[4];
            c[0] = tl[cc * 4 + 0], c[1] = tl[cc * 4 + 1], c[2] = tl[cc * 4 + 2], c[3] = tl[cc * 4 + 3];
            uint32_t key[Q];
            bool any = false;
#pragma unroll
            for (int j = 0; j < Q; j++)
            {
                key[j] = hamming_key<F>(q[j], c, k0 + cc);
                any |= key[j] < m2[j];
            }
            if (__any_sync(0xFFFFFFFFu, any))
            {
#pragma unroll
                for (int j = 0; j < Q; j++)
                {
                    const uint32_t hi = max(m1[j], key[j]);
                    m1[j] = min(m1[j], key[j]);
                    m2[j] = min(m2[j], hi);
                }
            }
        }
        __syncthreads(); // every warp is done with stage s
        if (tid == 0 && t + K1_STAGES < ntiles)
            issue(t + K1_STAGES);
    }

    const uint32_t splits = pp->splits;
#pragma unroll
    for (int j = 0; j < Q; j++)
    {
        const uint32_t qi = q_base + j * K1_THREADS + tid;
        if (qi < n_q)
        {
            if (splits == 1)
                pp->out[qi] = finish_keys(m1[j], m2[j], 0);
            else
                pp->partial[(size_t)split * n_q + qi] = make_uint2(m1[j], m2[j]);
        }
    }
}

// Top-2 of the union of the per-split (m1,m2) keys, in candidate-position order.
__global__ void __launch_bounds__(256)
    k1_merge_kernel(const __grid_constant__ K1Inline inl, const K1Problem *__restrict__ problems, uint32_t n_problems,
                    const uint32_t *__restrict__ merge_begin_g)
{
    const uint32_t *merge_begin = problems ? merge_begin_g : inl.merge_begin;
    if (!problems)
        problems = inl.p;
    // merge_begin[p] = first global query slot of problem p (prefix sum of n_q over split problems, 0 for
    // unsplit ones which own no slots); merge_begin[n_problems] = total
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= merge_begin[n_problems])
        return;
    uint32_t lo = 0, hi = n_problems - 1;
    while (lo < hi)
    {
        const uint32_t mid = (lo + hi + 1) >> 1;
        if (merge_begin[mid] <= g)
            lo = mid;
        else
            hi = mid - 1;
    }
    const K1Problem *pp = problems + lo;
    const uint32_t qi = g - merge_begin[lo];
    const uint32_t n_q = pp->n_q, splits = pp->splits;
    const uint32_t span = pp->tiles_per_split * K1_TILE_C;
    uint64_t a = ~0ull, b = ~0ull; // two smallest (distance, global position) keys
    for (uint32_t s = 0; s < splits; s++)
    {
        const uint2 m = pp->partial[(size_t)s * n_q + qi];
#pragma unroll
        for (int w = 0; w < 2; w++)
        {
            const uint32_t k = w == 0 ? m.x : m.y;
            if (k == 0xFFFFFFFFu)
                continue;
            const uint64_t key = ((uint64_t)(k >> K1_KEY_SHIFT) << 32) | (uint64_t)(s * span + (k & K1_KEY_IDX_MASK));
            const uint64_t hi2 = key > a ? key : a;
            a = key < a ? key : a;
            b = hi2 < b ? hi2 : b;
        }
    }
    ocb_top2 r;
    if (a == ~0ull)
    {
        r.best_k = 0;
        r.best_d = OCB_DIST_INF;
    }
    else
    {
        r.best_k = (uint32_t)a;
        r.best_d = (uint16_t)(a >> 32);
    }
    r.second_d = b == ~0ull ? (uint16_t)OCB_DIST_INF : (uint16_t)(b >> 32);
    pp->out[qi] = r;
}

// col_best_q[j] = best_k of row j's own top-2 (second pass of the cross-check), OCB_NO_INDEX if nothing was seen
__global__ void __launch_bounds__(256)
    k1_extract_best_kernel(const ocb_top2 *__restrict__ top2, uint32_t n, bool empty, uint32_t *__restrict__ best)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        best[i] = (empty || top2[i].best_d == OCB_DIST_INF) ? OCB_NO_INDEX : top2[i].best_k;
}

int k1_extract_best(const ocb_top2 *d_top2, uint32_t n, bool empty, uint32_t *d_best, cudaStream_t stream)
{
    if (n == 0)
        return 0;
    k1_extract_best_kernel<<<(n + 255) / 256, 256, 0, stream>>>(d_top2, n, empty, d_best);
    count_launch();
    OCB_CUDA(cudaGetLastError());
    return 0;
}

// ----------------------------------------------------------------------------------------------------------
// variants + host-side planning / launch
// ----------------------------------------------------------------------------------------------------------
struct K1Variant
{
    int q, f;
    void (*kernel)(const K1Inline, const K1Problem *, uint32_t);
    const char *name;
};
#define K1V(Q_, F_, MINB_)                                                                                             \
    {                                                                                                                  \
        Q_, F_, k1_top2_kernel<Q_, F_, MINB_>, "q" #Q_ "f" #F_                                                         \
    }
static const K1Variant k1_variants[] = {
    K1V(4, 7, 4),  // 0: default
    K1V(4, 0, 4),  // 1: plain 16-POPC form (the "naive POPC roofline" shape)
    K1V(4, 5, 4),  // 2
    K1V(4, 6, 4),  // 3
    K1V(4, 8, 4),  // 4
    K1V(4, 9, 4),  // 5
    K1V(4, 11, 4), // 6
    K1V(2, 7, 6),  // 7
    K1V(2, 8, 6),  // 8
    K1V(3, 7, 5),  // 9
    K1V(3, 8, 5),  // 10
    K1V(2, 0, 6),  // 11
};
constexpr int K1_NUM_VARIANTS = sizeof(k1_variants) / sizeof(k1_variants[0]);

static const K1Variant &current_variant()
{
    int v = options().k1_variant;
    if (v < 0 || v >= K1_NUM_VARIANTS)
        v = 0;
    return k1_variants[v];
}

int k1_queries_per_cta()
{
    return current_variant().q * K1_THREADS;
}

K1Plan k1_plan(K1Problem *problems, size_t n, size_t *partial_elems, int sms)
{
    K1Plan plan;
    const uint32_t tq = (uint32_t)k1_queries_per_cta();
    uint64_t base_items = 0;
    for (size_t p = 0; p < n; p++)
    {
        problems[p].q_tiles = (problems[p].n_q + tq - 1) / tq;
        base_items += problems[p].q_tiles;
    }
    // Split the candidate axis only as far as needed to give every SM `k1_items_per_sm` work items.
    const uint64_t target = (uint64_t)sms * (uint64_t)std::max(1, options().k1_items_per_sm);
    const uint32_t want_splits = base_items == 0 ? 1 : (uint32_t)std::min<uint64_t>((target + base_items - 1) / base_items, 1u << 16);
    uint32_t item = 0;
    for (size_t p = 0; p < n; p++)
    {
        K1Problem &P = problems[p];
        const uint32_t ctiles = (P.n_c + K1_TILE_C - 1) / K1_TILE_C;
        uint32_t splits = std::max(1u, std::min(want_splits, ctiles));
        uint32_t tps = ctiles == 0 ? 1 : (ctiles + splits - 1) / splits;
        if (tps > K1_MAX_TILES_PER_SPLIT - 1)
            tps = K1_MAX_TILES_PER_SPLIT - 1;
        splits = ctiles == 0 ? 1 : (ctiles + tps - 1) / tps;
        P.splits = splits;
        P.tiles_per_split = tps;
        P.item_begin = item;
        item += P.q_tiles * splits;
        partial_elems[p] = splits > 1 ? (size_t)splits * P.n_q : 0;
        plan.any_split |= splits > 1;
    }
    plan.total_items = item;
    return plan;
}

void k1_merge_begin(const K1Problem *problems, size_t n, uint32_t *merge_begin)
{
    uint32_t acc = 0;
    for (size_t p = 0; p < n; p++)
    {
        merge_begin[p] = acc;
        acc += problems[p].splits > 1 ? problems[p].n_q : 0;
    }
    merge_begin[n] = acc;
}

int k1_launch(const K1Problem *d_problems, const K1Problem *h_problems, size_t n, const K1Plan &plan,
              cudaStream_t stream)
{
    if (plan.total_items == 0)
        return 0;
    const K1Variant &v = current_variant();
    K1Inline inl;
    memset(&inl, 0, sizeof inl);
    const bool use_inline = n <= (size_t)K1_INLINE;
    if (use_inline)
    {
        for (size_t p = 0; p < n; p++)
            inl.p[p] = h_problems[p];
        k1_merge_begin(h_problems, n, inl.merge_begin);
        d_problems = nullptr;
    }
    v.kernel<<<plan.total_items, K1_THREADS, 0, stream>>>(inl, d_problems, (uint32_t)n);
    count_launch();
    OCB_CUDA(cudaGetLastError());
    if (plan.any_split)
    {
        // for uploaded tables merge_begin lives right behind the n problems (the C-ABI layer lays it out so)
        const uint32_t *d_merge_begin = use_inline ? nullptr : reinterpret_cast<const uint32_t *>(d_problems + n);
        uint64_t total = 0;
        for (size_t p = 0; p < n; p++)
            total += h_problems[p].splits > 1 ? h_problems[p].n_q : 0;
        if (total > 0)
        {
            const uint32_t blocks = (uint32_t)((total + 255) / 256);
            k1_merge_kernel<<<blocks, 256, 0, stream>>>(inl, d_problems, (uint32_t)n, d_merge_begin);
            count_launch();
            OCB_CUDA(cudaGetLastError());
        }
    }
    return 0;
}

} // namespace ocb
