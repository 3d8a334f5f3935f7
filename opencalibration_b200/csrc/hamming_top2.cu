// K1 -- brute-force Hamming top-2 over 512-bit descriptor rows (sm_100a), one kernel launch per submission.
//
// Replaces the loop nest of match_features_subset (reference src/match/match_features.cpp:71-93):
// for every query row the first candidate position at minimum Hamming distance, that distance, and the
// second-smallest distance counted with multiplicity (tie rule of :80-92). Optionally, in the same sweep, the
// column-wise best query of every candidate (cross-check; not in the reference).
//
// Mapping to the hardware
//   * one work item = one CTA = (query tile of 128*Q rows) x (a contiguous range of candidate rows);
//     every thread keeps Q query rows in registers (16 x u32 each, loaded as 128-bit vectors);
//   * candidate rows stream through shared memory in TILE_C-row tiles, copied by 1-D bulk TMA
//     (cp.async.bulk + mbarrier, SASS UBLKCP) into a K1_STAGES-deep ring, and are read back as warp-uniform
//     LDS.128 broadcasts, so one shared-memory read feeds 32 lanes x Q comparisons;
//   * distance = XOR + POPC on the integer pipes. POPC issues on the quarter-rate XU pipe (16/clk/SM), LOP3 on
//     the ALU pipe (64/clk/SM) and IMAD on the FMA pipe (64/clk/SM), and the three overlap, so part of the 16
//     popcounts per comparison is traded for carry-save adders (F full adders = 2 LOP3 each, each removes one
//     POPC) and the weighted sum of the remaining popcounts is accumulated with IMADs (runtime multiplier, so
//     ptxas cannot turn them back into ALU shifts/adds) into a packed value
//           v = (distance << 20) | low bits
//     The low bits are the candidate position (branch-free update: within one query v then orders by (distance,
//     position), i.e. the first-seen minimum wins like the reference's strict-less-than updates) or the query's
//     slot in the tile (vote form and cross-check: across the queries of a column v orders by (distance, query
//     position));
//   * per query the two smallest values (s1 <= s2) are kept branch-free (two min, one max) with the candidate
//     position in the low bits of v; the alternative - one compare per comparison and a warp-uniform branch around
//     the rare update, the position kept apart - is selectable (k1_update) and lost at every run length measured;
//   * cross-check: per candidate the warp-wide minimum of v (REDUX) is parked in a double-buffered shared-memory
//     array; after every tile 64 threads combine the warps' minima and issue one 64-bit atomic max per candidate
//     and CTA on the complemented (distance, query) key in global memory (complemented so that the
//     zero-initialised workspace means "nothing yet");
//   * when a single pair cannot fill 148 SMs the candidate axis is split across CTAs; every CTA folds its
//     per-query (d1, position, d2) into a global per-query state with two atomics (64-bit max on the
//     complemented (distance, position) key; 32-bit max on the complemented distance of every key that loses) and
//     takes a ticket per query tile; the CTA that draws the last ticket converts the state to the final records.
//     A second ticket per candidate split lets the last CTA of a split convert the column keys to query
//     indices. No second kernel, no serial merge.
#include "ocb_internal.cuh"

#include <algorithm>
#include <cstring>
#include <vector>

namespace ocb
{

constexpr int K1_THREADS = 128;
constexpr int K1_TILE_C = 64; // candidate rows per shared-memory tile (4 KB)
constexpr int K1_STAGES = 4;
constexpr int K1_SHIFT = 20;  // v = distance << 20 | slot ; candidates per work item < 2^20 (index kept apart)
constexpr uint32_t K1_SLOT_MASK = (1u << K1_SHIFT) - 1;
constexpr uint32_t K1_MIN_ROWS_PER_SPLIT = 32; // do not split finer than this
constexpr uint32_t K1_NONE = 0xFFFFFFFFu;

__device__ __forceinline__ uint32_t mad_u32(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// ----------------------------------------------------------------------------------------------------------
// v = v0 + (popcount(q ^ c) << 20), with F carry-save full adders in front of the popcounts.
// IMADACC: accumulate with IMAD on the FMA pipe (w = 1 << 20 arrives as a kernel parameter).
// ----------------------------------------------------------------------------------------------------------
template <int F, bool IMADACC>
__device__ __forceinline__ uint32_t hamming_value(const uint32_t (&q)[16], const uint4 (&c)[4], uint32_t v0, uint32_t w)
{
    constexpr int F1 = F < 7 ? F : 7;                        // adders on weight-1 words
    constexpr int F2 = F <= 7 ? 0 : (F - 7 < 3 ? F - 7 : 3); // adders on weight-2 words
    constexpr int F4 = F <= 10 ? 0 : 1;                      // adder on weight-4 words
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
    uint32_t v = v0;
    if constexpr (IMADACC)
    {
        const uint32_t wa = w, wb = w * 2, wc = w * 4, wd = w * 8;
#pragma unroll
        for (int i = h1; i < t1; i++)
            v = mad_u32((uint32_t)__popc(w1[i]), wa, v);
#pragma unroll
        for (int i = h2; i < t2; i++)
            v = mad_u32((uint32_t)__popc(w2[i]), wb, v);
#pragma unroll
        for (int i = h4; i < t4; i++)
            v = mad_u32((uint32_t)__popc(w4[i]), wc, v);
#pragma unroll
        for (int i = 0; i < t8; i++)
            v = mad_u32((uint32_t)__popc(w8[i]), wd, v);
    }
    else
    {
#pragma unroll
        for (int i = h1; i < t1; i++)
            v += (uint32_t)__popc(w1[i]) * (1u << K1_SHIFT);
#pragma unroll
        for (int i = h2; i < t2; i++)
            v += (uint32_t)__popc(w2[i]) * (2u << K1_SHIFT);
#pragma unroll
        for (int i = h4; i < t4; i++)
            v += (uint32_t)__popc(w4[i]) * (4u << K1_SHIFT);
#pragma unroll
        for (int i = 0; i < t8; i++)
            v += (uint32_t)__popc(w8[i]) * (8u << K1_SHIFT);
    }
    return v;
}

// ----------------------------------------------------------------------------------------------------------
// Prefix form. Both rows are first mapped by the same invertible GF(2)-linear map T (prefix_transform below):
//     T(r)[2k] = r[0] ^ r[1] ^ ... ^ r[2k]   (k = 1..7),  every other word unchanged.
// XOR commutes with T, so with x = q ^ c:  T(q)[2k] ^ T(c)[2k] = x[0] ^ ... ^ x[2k] =: s_k, i.e. the SUM output of a
// chain of 7 full adders (x0,x1,x2), (s_1,x3,x4), ..., (s_6,x13,x14) costs ONE XOR each instead of three XORs and
// a 3-input XOR. The CARRY of adder k needs its three inputs, but the third is implied by the sum:
//     maj(a, b, a^b^s) = (a & b) | ((a ^ b) & ~s)            -> one LOP3 (0xD4) on (a, b, s)
// Level 1 therefore costs 9 + 7 + 7 = 23 LOP3 for 16 words (16 + 14 = 30 in the plain form) and leaves two
// weight-1 words (s_7, x15) and seven weight-2 words; G further adders reduce the weight-2/4 words as before.
// ----------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t lop3_carry_implied(uint32_t a, uint32_t b, uint32_t s)
{
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xD4;" : "=r"(d) : "r"(a), "r"(b), "r"(s));
    return d;
}

__device__ __forceinline__ void prefix_transform(uint32_t (&r)[16])
{
    r[2] = lop3_xor3(r[0], r[1], r[2]);
#pragma unroll
    for (int k = 2; k <= 7; k++)
        r[2 * k] = lop3_xor3(r[2 * k - 2], r[2 * k - 1], r[2 * k]);
}

template <int G, bool IMADACC>
__device__ __forceinline__ uint32_t hamming_value_pfx(const uint32_t (&q)[16], const uint4 (&c4)[4], uint32_t v0, uint32_t w)
{
    constexpr int G2 = G < 3 ? G : 3;     // adders on weight-2 words
    constexpr int G4 = G <= 3 ? 0 : 1;    // adder on weight-4 words
    const uint32_t c[16] = {c4[0].x, c4[0].y, c4[0].z, c4[0].w, c4[1].x, c4[1].y, c4[1].z, c4[1].w,
                            c4[2].x, c4[2].y, c4[2].z, c4[2].w, c4[3].x, c4[3].y, c4[3].z, c4[3].w};
    uint32_t w2[7 + 3];
    uint32_t w4[3 + 1];
    uint32_t w8[1];
    uint32_t a = q[0] ^ c[0];
    uint32_t s = a; // running sum word
#pragma unroll
    for (int k = 1; k <= 7; k++)
    {
        const uint32_t b = q[2 * k - 1] ^ c[2 * k - 1];
        const uint32_t sn = q[2 * k] ^ c[2 * k];
        w2[k - 1] = lop3_carry_implied(s, b, sn);
        s = sn;
    }
    const uint32_t x15 = q[15] ^ c[15];
    int h2 = 0, t2 = 7, h4 = 0, t4 = 0, t8 = 0;
#pragma unroll
    for (int f = 0; f < G2; f++)
    {
        w2[t2++] = lop3_xor3(w2[h2], w2[h2 + 1], w2[h2 + 2]);
        w4[t4++] = lop3_maj(w2[h2], w2[h2 + 1], w2[h2 + 2]);
        h2 += 3;
    }
#pragma unroll
    for (int f = 0; f < G4; f++)
    {
        w4[t4++] = lop3_xor3(w4[h4], w4[h4 + 1], w4[h4 + 2]);
        w8[t8++] = lop3_maj(w4[h4], w4[h4 + 1], w4[h4 + 2]);
        h4 += 3;
    }
    uint32_t v = v0;
    if constexpr (IMADACC)
    {
        const uint32_t wa = w, wb = w * 2, wc = w * 4, wd = w * 8;
        v = mad_u32((uint32_t)__popc(s), wa, v);
        v = mad_u32((uint32_t)__popc(x15), wa, v);
#pragma unroll
        for (int i = h2; i < t2; i++)
            v = mad_u32((uint32_t)__popc(w2[i]), wb, v);
#pragma unroll
        for (int i = h4; i < t4; i++)
            v = mad_u32((uint32_t)__popc(w4[i]), wc, v);
#pragma unroll
        for (int i = 0; i < t8; i++)
            v = mad_u32((uint32_t)__popc(w8[i]), wd, v);
    }
    else
    {
        v += ((uint32_t)__popc(s) + (uint32_t)__popc(x15)) * (1u << K1_SHIFT);
#pragma unroll
        for (int i = h2; i < t2; i++)
            v += (uint32_t)__popc(w2[i]) * (2u << K1_SHIFT);
#pragma unroll
        for (int i = h4; i < t4; i++)
            v += (uint32_t)__popc(w4[i]) * (4u << K1_SHIFT);
#pragma unroll
        for (int i = 0; i < t8; i++)
            v += (uint32_t)__popc(w8[i]) * (8u << K1_SHIFT);
    }
    return v;
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

__device__ __forceinline__ uint32_t dist_of(uint32_t v)
{
    return v == K1_NONE ? (uint32_t)OCB_DIST_INF : (v >> K1_SHIFT);
}

__device__ __forceinline__ ocb_top2 make_record(uint32_t d1, uint32_t d2, uint32_t idx)
{
    ocb_top2 r;
    r.best_k = d1 == OCB_DIST_INF ? 0u : idx; // feature_match best_match{i, 0, inf} (match_features.cpp:74)
    r.best_d = (uint16_t)d1;
    r.second_d = (uint16_t)d2;
    return r;
}

template <int Q, int F, bool IMADACC, bool COL, int MINB, bool PFX, bool BF>
__global__ void __launch_bounds__(K1_THREADS, MINB)
    k1_top2_kernel(const __grid_constant__ K1Inline inl, const K1Problem *__restrict__ problems, uint32_t n_problems,
                   uint32_t w)
{
    static_assert(!PFX || F >= 7, "the prefix form starts from the 7-adder chain");
    __shared__ alignas(128) uint4 tile[K1_STAGES][K1_TILE_C * 4];
    __shared__ alignas(8) uint64_t full_bar[K1_STAGES];
    __shared__ uint32_t ticket[2];
    // cross-check: per-warp column minima of a tile, combined by the CTA after the tile (double-buffered)
    __shared__ uint32_t colmin[COL ? 2 : 1][K1_THREADS / 32][COL ? K1_TILE_C : 1];

    const uint32_t tid = threadIdx.x;
    // <= K1_INLINE problems travel in the kernel parameters (no table upload on the single-pair path)
    const K1Problem *pp = problems ? find_problem(problems, n_problems, blockIdx.x)
                                   : &inl.p[(n_problems > 1 && blockIdx.x >= inl.p[1].item_begin) ? 1 : 0];
    const uint32_t n_q = pp->n_q, n_c = pp->n_c, q_tiles = pp->q_tiles;
    const uint32_t local = blockIdx.x - pp->item_begin;
    const uint32_t split = local / q_tiles;
    const uint32_t qtile = local - split * q_tiles;
    const uint32_t rows_per_split = pp->rows_per_split;
    const uint32_t c_begin = split * rows_per_split;
    const uint32_t c_end = min(n_c, c_begin + rows_per_split);
    const uint32_t c_cnt = c_end > c_begin ? c_end - c_begin : 0;
    const uint32_t ntiles = (c_cnt + K1_TILE_C - 1) / K1_TILE_C;
    const uint4 *__restrict__ cand = pp->c + (size_t)c_begin * 4;
    unsigned long long *__restrict__ col64 = COL ? pp->col64 : nullptr;

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

    // query rows -> registers (128-bit loads; rows past n_q read row n_q-1: same distances as the real last row
    // but a larger slot, so they never win a column and are never written back)
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
        if constexpr (PFX)
            prefix_transform(q[j]);
    }
    uint32_t s1[Q], s2[Q], bi[Q]; // two smallest values, candidate position (within this item) of the first minimum
#pragma unroll
    for (int j = 0; j < Q; j++)
        s1[j] = s2[j] = K1_NONE, bi[j] = 0;

    // Prefix form: candidate tile t is mapped by T in place, one row per thread, while tile t-1 is still being
    // consumed by the other warps; the __syncthreads that retires a stage also publishes the next mapped tile.
    auto transform_tile = [&](uint32_t t) {
        const uint32_t rows = min((uint32_t)K1_TILE_C, c_cnt - t * K1_TILE_C);
        if (tid < rows)
        {
            const uint32_t s = t % K1_STAGES;
            mbar_wait(&full_bar[s], (t / K1_STAGES) & 1);
            uint4 *row = &tile[s][tid * 4];
            uint4 a = row[0], b = row[1], c = row[2], d = row[3];
            a.z = lop3_xor3(a.x, a.y, a.z);
            b.x = lop3_xor3(a.z, a.w, b.x);
            b.z = lop3_xor3(b.x, b.y, b.z);
            c.x = lop3_xor3(b.z, b.w, c.x);
            c.z = lop3_xor3(c.x, c.y, c.z);
            d.x = lop3_xor3(c.z, c.w, d.x);
            d.z = lop3_xor3(d.x, d.y, d.z);
            row[0] = a, row[1] = b, row[2] = c, row[3] = d;
            fence_proxy_async_smem(); // these generic-proxy writes precede the next bulk copy into this stage
        }
    };
    if constexpr (PFX)
    {
        if (ntiles)
            transform_tile(0);
        __syncthreads();
    }

    for (uint32_t t = 0; t < ntiles; t++)
    {
        const uint32_t s = t % K1_STAGES;
        if constexpr (!PFX)
            mbar_wait(&full_bar[s], (t / K1_STAGES) & 1);
        const uint32_t rows = min((uint32_t)K1_TILE_C, c_cnt - t * K1_TILE_C);
        const uint4 *__restrict__ tl = &tile[s][0];
        const uint32_t k0 = t * K1_TILE_C;
#pragma unroll(Q <= 2 ? 4 : 2)
        for (uint32_t cc = 0; cc < rows; cc++)
        {
            uint4 c[4];
            c[0] = tl[cc * 4 + 0], c[1] = tl[cc * 4 + 1], c[2] = tl[cc * 4 + 2], c[3] = tl[cc * 4 + 3];
            uint32_t v[Q];
            if constexpr (BF)
            {
                // Branch-free form for short candidate runs (a warp carries 32*Q independent minima, so with runs of
                // a few hundred rows some lane updates at almost every candidate and the vote never skips):
                // v = distance << 20 | position, so the two smallest values ARE (first minimum, second with
                // multiplicity) and the position rides along in s1.
                const uint32_t kidx = k0 + cc;
#pragma unroll
                for (int j = 0; j < Q; j++)
                {
                    if constexpr (PFX)
                        v[j] = hamming_value_pfx<F - 7, IMADACC>(q[j], c, kidx, w);
                    else
                        v[j] = hamming_value<F, IMADACC>(q[j], c, kidx, w);
                    s2[j] = min(s2[j], max(s1[j], v[j]));
                    s1[j] = min(s1[j], v[j]);
                }
                if constexpr (COL)
                {
                    const uint32_t nk = tid - kidx; // position bits -> query slot bits
#pragma unroll
                    for (int j = 0; j < Q; j++)
                        v[j] = v[j] + nk + (uint32_t)(j * K1_THREADS);
                }
            }
            else
            {
                bool any = false;
#pragma unroll
                for (int j = 0; j < Q; j++)
                {
                    if constexpr (PFX)
                        v[j] = hamming_value_pfx<F - 7, IMADACC>(q[j], c, (uint32_t)(j * K1_THREADS) + tid, w);
                    else
                        v[j] = hamming_value<F, IMADACC>(q[j], c, (uint32_t)(j * K1_THREADS) + tid, w);
                    any |= v[j] < s2[j];
                }
                if (__any_sync(0xFFFFFFFFu, any))
                {
#pragma unroll
                    for (int j = 0; j < Q; j++)
                    {
                        // match_features.cpp:80-92 on values that differ only by distance within one query
                        const bool better = v[j] < s1[j];
                        s2[j] = min(s2[j], max(s1[j], v[j]));
                        s1[j] = min(s1[j], v[j]);
                        bi[j] = better ? k0 + cc : bi[j];
                    }
                }
            }
            if constexpr (COL)
            {
                uint32_t cm = v[0];
#pragma unroll
                for (int j = 1; j < Q; j++)
                    cm = min(cm, v[j]);
                colmin[t & 1][tid >> 5][cc] = __reduce_min_sync(0xFFFFFFFFu, cm); // every lane stores the same word
            }
        }
        if constexpr (PFX)
        {
            if (t + 1 < ntiles)
                transform_tile(t + 1);
        }
        __syncthreads(); // every warp is done with stage s (and, prefix form, tile t+1 is mapped)
        if (tid == 0 && t + K1_STAGES < ntiles)
            issue(t + K1_STAGES);
        if constexpr (COL)
        {
            // one atomic per candidate and CTA: thread k combines the warps' minima of candidate k0 + k. The other
            // warps are already in tile t+1 (other colmin buffer); buffer t&1 is written again in tile t+2, i.e.
            // after the next __syncthreads, which this thread reaches after this flush.
            if (tid < rows)
            {
                uint32_t cm = colmin[t & 1][0][tid];
#pragma unroll
                for (int wp = 1; wp < K1_THREADS / 32; wp++)
                    cm = min(cm, colmin[t & 1][wp][tid]);
                const unsigned long long key =
                    ((unsigned long long)(cm >> K1_SHIFT) << 32) | (unsigned long long)(q_base + (cm & K1_SLOT_MASK));
                red_max_u64_global(&col64[c_begin + k0 + tid], ~key);
            }
        }
    }

    if constexpr (BF)
    {
#pragma unroll
        for (int j = 0; j < Q; j++)
            bi[j] = s1[j] & K1_SLOT_MASK;
    }
    const uint32_t splits = pp->splits;
    if (splits == 1)
    {
#pragma unroll
        for (int j = 0; j < Q; j++)
        {
            const uint32_t qi = q_base + j * K1_THREADS + tid;
            if (qi < n_q)
                pp->out[qi] = make_record(dist_of(s1[j]), dist_of(s2[j]), bi[j]);
        }
        if (!COL)
            return;
    }
    else
    {
        // Merge this split into the per-query global state with atomics (no serial tail):
        //   best64[q] = max over splits of ~((d1 << 32) | position)   == the smallest (distance, position) key
        //   sec32[q]  = max of ~d over every key that is not the final best: each split's d2, each split's d1 that
        //               loses on arrival, and each former best at the moment it is displaced.
        unsigned long long *__restrict__ best64 = pp->best64;
        uint32_t *__restrict__ sec32 = pp->sec32;
        unsigned long long old[Q];
#pragma unroll
        for (int j = 0; j < Q; j++)
        {
            const uint32_t qi = q_base + j * K1_THREADS + tid;
            old[j] = 0;
            if (qi < n_q && s1[j] != K1_NONE)
            {
                const unsigned long long key =
                    ((unsigned long long)(s1[j] >> K1_SHIFT) << 32) | (unsigned long long)(c_begin + bi[j]);
                old[j] = atom_max_u64_global(&best64[qi], ~key);
            }
        }
#pragma unroll
        for (int j = 0; j < Q; j++)
        {
            const uint32_t qi = q_base + j * K1_THREADS + tid;
            if (qi < n_q && s1[j] != K1_NONE)
            {
                const unsigned long long key =
                    ((unsigned long long)(s1[j] >> K1_SHIFT) << 32) | (unsigned long long)(c_begin + bi[j]);
                const unsigned long long oldkey = ~old[j];
                uint32_t loser = K1_NONE; // distance that becomes a second-best candidate
                if (oldkey < key)
                    loser = s1[j] >> K1_SHIFT; // an earlier split already holds a better (distance, position)
                else if (old[j] != 0)
                    loser = (uint32_t)(oldkey >> 32); // we displaced the former best
                if (s2[j] != K1_NONE)
                    loser = min(loser, s2[j] >> K1_SHIFT);
                if (loser != K1_NONE)
                    red_max_u32_global(&sec32[qi], ~loser);
            }
        }
    }

    // ---- tickets: last CTA of a query tile writes its records, last CTA of a split converts its columns ----
    __threadfence();
    __syncthreads();
    if (tid == 0)
    {
        ticket[0] = splits > 1 ? atomicAdd(&pp->counters[qtile], 1u) : 0u;
        ticket[1] = COL ? atomicAdd(&pp->counters[q_tiles + split], 1u) : 0u;
    }
    __syncthreads();
    if (splits > 1 && ticket[0] == splits - 1)
    {
        __threadfence();
#pragma unroll
        for (int j = 0; j < Q; j++)
        {
            const uint32_t qi = q_base + j * K1_THREADS + tid;
            if (qi < n_q)
            {
                const unsigned long long b = ~__ldcg(&pp->best64[qi]); // ~0 (none) when the state is still zero
                const uint32_t sd = ~__ldcg(&pp->sec32[qi]);
                const uint32_t d1 = b == ~0ull ? (uint32_t)OCB_DIST_INF : (uint32_t)(b >> 32);
                pp->out[qi] = make_record(d1, sd == K1_NONE ? (uint32_t)OCB_DIST_INF : sd, (uint32_t)b);
            }
        }
    }
    if (COL && ticket[1] == q_tiles - 1)
    {
        __threadfence();
        uint32_t *__restrict__ col_out = pp->col_out;
        for (uint32_t k = c_begin + tid; k < c_end; k += K1_THREADS)
            col_out[k] = (uint32_t)(~__ldcg(&col64[k])); // zero (nothing seen) -> OCB_NO_INDEX
    }
}

// ----------------------------------------------------------------------------------------------------------
// variants + host-side planning / launch
// ----------------------------------------------------------------------------------------------------------
typedef void (*K1Kernel)(const K1Inline, const K1Problem *, uint32_t, uint32_t);
struct K1Variant
{
    int q, f;
    K1Kernel kern[2][2]; // [branch-free update][cross-check]
    const char *name;
};
#define K1V(Q_, F_, A_, MINB_, P_)                                                                                     \
    {                                                                                                                  \
        Q_, F_,                                                                                                        \
            {{k1_top2_kernel<Q_, F_, A_, false, MINB_, P_, false>, k1_top2_kernel<Q_, F_, A_, true, MINB_, P_, false>},\
             {k1_top2_kernel<Q_, F_, A_, false, MINB_, P_, true>, k1_top2_kernel<Q_, F_, A_, true, MINB_, P_, true>}}, \
            "q" #Q_ "f" #F_ "a" #A_ "p" #P_                                                                            \
    }
// FC_ = adders of the cross-check instantiations: the cross-check's own work sits on the ALU pipe, so its best
// balance has fewer adders (more POPCs on the XU pipe) than the plain sweep's
#define K1V2(Q_, F_, FC_, A_, MINB_, P_)                                                                               \
    {                                                                                                                  \
        Q_, F_,                                                                                                        \
            {{k1_top2_kernel<Q_, F_, A_, false, MINB_, P_, false>, k1_top2_kernel<Q_, FC_, A_, true, MINB_, P_, false>},\
             {k1_top2_kernel<Q_, F_, A_, false, MINB_, P_, true>, k1_top2_kernel<Q_, FC_, A_, true, MINB_, P_, true>}}, \
            "q" #Q_ "f" #F_ "c" #FC_ "a" #A_ "p" #P_                                                                   \
    }
static const K1Variant k1_variants[] = {
    K1V2(2, 9, 7, true, 6, true), // 0: default: prefix form, two queries per thread so that six CTAs (24 warps) are
                                  //    resident per SM; 9 adders (27 LOP3 + 7 POPC + 7 IMAD per comparison) for the
                                  //    plain sweep, 7 (23 LOP3 + 9 POPC + 9 IMAD) with the fused cross-check
    K1V(4, 0, false, 4, false),  // 1: plain 16-POPC form (the "naive POPC roofline" shape)
    K1V(4, 7, true, 4, false),   // 2: first-round default: plain carry-save form, 7 adders (30 LOP3 + 9 POPC)
    K1V(4, 8, true, 4, true),    // 3
    K1V(4, 10, true, 4, true),   // 4
    K1V(4, 9, false, 4, true),   // 5: compiler-chosen accumulation (IADD3/LEA on the ALU pipe)
    K1V(3, 9, true, 4, true),    // 6
    K1V(3, 9, true, 5, true),    // 7
    K1V(4, 9, true, 4, true),    // 8: four queries per thread (half the shared-memory reads, 16 warps per SM)
    K1V(4, 11, true, 4, true),   // 9
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

static int resident_ctas_per_sm(const K1Variant &v)
{
    static int cache[K1_NUM_VARIANTS] = {0};
    const int idx = (int)(&v - k1_variants);
    if (cache[idx] == 0)
    {
        int n = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, v.kern[1][1], K1_THREADS, 0) != cudaSuccess || n <= 0)
        {
            cudaGetLastError();
            n = 4;
        }
        cache[idx] = n;
    }
    return cache[idx];
}

K1Plan k1_plan(K1Problem *problems, size_t n, int sms, int items_per_sm)
{
    K1Plan plan;
    const K1Variant &v = current_variant();
    const uint32_t tq = (uint32_t)(v.q * K1_THREADS);
    uint64_t base_items = 0;
    for (size_t p = 0; p < n; p++)
    {
        problems[p].q_tiles = (problems[p].n_q + tq - 1) / tq;
        base_items += problems[p].q_tiles;
    }
    // Split the candidate axis (at row granularity) only as far as needed to give every SM ~items_per_sm work
    // items, and so that the equal-sized items fill a whole number of waves of resident CTAs (no ragged tail).
    if (items_per_sm <= 0)
        items_per_sm = std::max(1, options().k1_items_per_sm);
    uint64_t target = (uint64_t)sms * (uint64_t)items_per_sm;
    const uint64_t slots = (uint64_t)sms * (uint64_t)(items_per_sm >= (1 << 20) ? 1 : resident_ctas_per_sm(v));
    if (target >= slots)
        target = target / slots * slots;
    const uint32_t want_splits =
        base_items == 0 ? 1 : (uint32_t)std::min<uint64_t>(std::max<uint64_t>(target / base_items, 1), 1u << 16);
    uint32_t item = 0;
    for (size_t p = 0; p < n; p++)
    {
        K1Problem &P = problems[p];
        uint32_t rps = P.n_c == 0 ? 1 : (P.n_c + want_splits - 1) / want_splits;
        rps = std::max(rps, std::min<uint32_t>(P.n_c, K1_MIN_ROWS_PER_SPLIT));
        rps = std::max(1u, std::min(rps, (1u << K1_SHIFT) - 1));
        const uint32_t splits = P.n_c == 0 ? 1 : (P.n_c + rps - 1) / rps;
        P.splits = splits;
        P.rows_per_split = rps;
        P.item_begin = item;
        item += P.q_tiles * splits;
        plan.any_col |= P.col_out != nullptr;
    }
    plan.total_items = item;
    return plan;
}

size_t k1_state_bytes(const K1Problem &P)
{
    // [tickets: q_tiles + splits u32][best64: n_q u64][sec32: n_q u32][col64: n_c u64], each 8-byte aligned
    size_t b = 0;
    if (P.splits > 1 || P.col_out)
        b += ((size_t)P.q_tiles + P.splits + 1) / 2 * 8;
    if (P.splits > 1)
        b += (size_t)P.n_q * 8 + ((size_t)P.n_q + 1) / 2 * 8;
    if (P.col_out)
        b += (size_t)P.n_c * 8;
    return b;
}
void k1_bind_state(K1Problem &P, void *d_state)
{
    char *p = static_cast<char *>(d_state);
    P.counters = reinterpret_cast<uint32_t *>(p);
    if (P.splits > 1 || P.col_out)
        p += ((size_t)P.q_tiles + P.splits + 1) / 2 * 8;
    P.best64 = reinterpret_cast<unsigned long long *>(p);
    P.sec32 = reinterpret_cast<uint32_t *>(p + (P.splits > 1 ? (size_t)P.n_q * 8 : 0));
    if (P.splits > 1)
        p += (size_t)P.n_q * 8 + ((size_t)P.n_q + 1) / 2 * 8;
    P.col64 = P.col_out ? reinterpret_cast<unsigned long long *>(p) : nullptr;
}

int k1_launch(const K1Problem *d_problems, const K1Problem *h_problems, size_t n, const K1Plan &plan,
              cudaStream_t stream)
{
    if (plan.total_items == 0)
        return 0;
    const K1Variant &v = current_variant();
    K1Inline inl;
    memset(&inl, 0, sizeof inl);
    if (n <= (size_t)K1_INLINE)
    {
        for (size_t p = 0; p < n; p++)
            inl.p[p] = h_problems[p];
        d_problems = nullptr;
    }
    // update form: branch-free while the candidate runs are short (every candidate then updates some lane of a warp),
    // vote-and-skip when they are long (updates become rare). option k1_update: 0 = by run length, 1 = vote, 2 = branch-free
    uint32_t min_run = 0xFFFFFFFFu;
    for (size_t p = 0; p < n; p++)
        if (h_problems[p].n_c)
            min_run = std::min(min_run, h_problems[p].rows_per_split);
    const int upd = options().k1_update;
    const bool bf = upd == 2 || (upd == 0 && min_run < (uint32_t)std::max(1, options().k1_bf_rows));
    K1Kernel k = v.kern[bf ? 1 : 0][plan.any_col ? 1 : 0];
    k<<<plan.total_items, K1_THREADS, 0, stream>>>(inl, d_problems, (uint32_t)n, 1u << K1_SHIFT);
    count_launch();
    OCB_CUDA(cudaGetLastError());
    return 0;
}

} // namespace ocb
