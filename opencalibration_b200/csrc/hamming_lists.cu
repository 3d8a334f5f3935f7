// K4 -- Hamming top-2 over per-query CANDIDATE LISTS (sm_100a): the guided matcher of the dense stage.
//
// Replaces the inner loop of densifyMesh (reference src/dense/dense_stereo.cpp:251-273): for one source feature and
// the features `nearby` its predicted position in a candidate image (KD-tree radius search, :244-246), the first
// list position at minimum Hamming distance, that distance, and the second-smallest distance with multiplicity --
// the same update rule as match_features_subset (src/match/match_features.cpp:80-92), on an irregular list instead
// of a dense candidate range. The acceptance rule (:275-276: ratio 0.85 with >= 2 candidates, absolute 0.35 with
// one) stays in the C++ adapter, in double, like the ratio test of K1.
//
// Bound: memory, not the integer pipes. Every comparison gathers one 64-byte candidate row through an index, so the
// kernel moves 64 B (+4 B of index) per 16 XOR + POPC; at the 16 POPC/clk/SM rate the pipes could take ~290 G
// comparisons/s while HBM delivers 6.5 TB/s / 68 B = 96 G rows/s (more from L2: neighbouring queries share most of
// their candidates, and the reference walks the queries in Hilbert order, :31-48,191-193). The work is therefore laid
// out for the memory system:
//   * one list per GROUP of 8 lanes (4 lists per warp): lists are short (~100 rows at the reference's 150 px radius)
//     and a full warp per list would idle most lanes on the tail and spend 5 shuffle rounds per merge;
//   * each lane of a group fetches whole candidate rows -- 4 x LDG.128 per row, two rows in flight per lane -- so a
//     group has 16 independent 128-bit loads outstanding and every fetched 32-byte sector is fully used;
//   * the query row is loaded once per list (4 x LDG.128, broadcast within the group) and kept in registers;
//   * key = (distance << 22) | list position: the two smallest keys of a list ARE (first minimum, second with
//     multiplicity); lanes keep their two smallest keys branch-free and the group merges them with 3 shuffle rounds.
#include "ocb_internal.cuh"

namespace ocb
{

constexpr int K4_THREADS = 256;
constexpr int K4_GROUP = 8;                               // lanes per list
constexpr int K4_LISTS_PER_CTA = K4_THREADS / K4_GROUP;   // 32
constexpr int K4_POS_BITS = 22;                           // list length < 2^22 (checked by the C entry points)
constexpr uint32_t K4_POS_MASK = (1u << K4_POS_BITS) - 1;
constexpr uint32_t K4_NONE = 0xFFFFFFFFu;

__device__ __forceinline__ uint32_t k4_distance(const uint4 (&q)[4], const uint4 (&c)[4])
{
    // 16 XOR, then 4 full adders on (x0,x1,x2) ... (x9,x10,x11): 12 words -> 4 sums + 4 carries, 12 POPC in all
    uint32_t x[16] = {q[0].x ^ c[0].x, q[0].y ^ c[0].y, q[0].z ^ c[0].z, q[0].w ^ c[0].w,
                      q[1].x ^ c[1].x, q[1].y ^ c[1].y, q[1].z ^ c[1].z, q[1].w ^ c[1].w,
                      q[2].x ^ c[2].x, q[2].y ^ c[2].y, q[2].z ^ c[2].z, q[2].w ^ c[2].w,
                      q[3].x ^ c[3].x, q[3].y ^ c[3].y, q[3].z ^ c[3].z, q[3].w ^ c[3].w};
    uint32_t ones = 0, twos = 0;
#pragma unroll
    for (int f = 0; f < 4; f++)
    {
        ones += __popc(lop3_xor3(x[3 * f], x[3 * f + 1], x[3 * f + 2]));
        twos += __popc(lop3_maj(x[3 * f], x[3 * f + 1], x[3 * f + 2]));
    }
#pragma unroll
    for (int i = 12; i < 16; i++)
        ones += __popc(x[i]);
    return ones + 2 * twos;
}

__device__ __forceinline__ void k4_push(uint32_t &s1, uint32_t &s2, uint32_t key)
{
    s2 = min(s2, max(s1, key));
    s1 = min(s1, key);
}

__global__ void __launch_bounds__(K4_THREADS)
    k4_lists_kernel(const uint4 *__restrict__ q_rows, const uint4 *__restrict__ c_rows,
                    const uint32_t *__restrict__ list_query, const uint64_t *__restrict__ list_begin,
                    const uint32_t *__restrict__ list_candidates, uint32_t n_lists, ocb_top2 *__restrict__ out)
{
    const uint32_t tid = threadIdx.x;
    const uint32_t sub = tid & (K4_GROUP - 1);
    const uint32_t list = blockIdx.x * K4_LISTS_PER_CTA + tid / K4_GROUP;
    // whole groups leave together (the shuffles below only span a group, and the mask names only its lanes)
    if (list >= n_lists)
        return;
    const uint32_t gmask = 0xFFu << ((tid & 31) & ~(K4_GROUP - 1));

    const uint64_t begin = list_begin[list];
    const uint32_t len = (uint32_t)(list_begin[list + 1] - begin);
    const uint32_t *__restrict__ cand = list_candidates + begin;
    uint4 q[4];
    {
        const uint4 *row = q_rows + (size_t)list_query[list] * 4;
        q[0] = __ldg(row + 0), q[1] = __ldg(row + 1), q[2] = __ldg(row + 2), q[3] = __ldg(row + 3);
    }
    uint32_t s1 = K4_NONE, s2 = K4_NONE;
    uint32_t i = sub;
    for (; i + K4_GROUP < len; i += 2 * K4_GROUP) // two rows in flight per lane
    {
        const uint4 *ra = c_rows + (size_t)__ldg(cand + i) * 4;
        const uint4 *rb = c_rows + (size_t)__ldg(cand + i + K4_GROUP) * 4;
        uint4 a[4], b[4];
        a[0] = __ldg(ra + 0), a[1] = __ldg(ra + 1), a[2] = __ldg(ra + 2), a[3] = __ldg(ra + 3);
        b[0] = __ldg(rb + 0), b[1] = __ldg(rb + 1), b[2] = __ldg(rb + 2), b[3] = __ldg(rb + 3);
        k4_push(s1, s2, (k4_distance(q, a) << K4_POS_BITS) | i);
        k4_push(s1, s2, (k4_distance(q, b) << K4_POS_BITS) | (i + K4_GROUP));
    }
    if (i < len)
    {
        const uint4 *ra = c_rows + (size_t)__ldg(cand + i) * 4;
        uint4 a[4];
        a[0] = __ldg(ra + 0), a[1] = __ldg(ra + 1), a[2] = __ldg(ra + 2), a[3] = __ldg(ra + 3);
        k4_push(s1, s2, (k4_distance(q, a) << K4_POS_BITS) | i);
    }
#pragma unroll
    for (int o = K4_GROUP / 2; o >= 1; o >>= 1)
    {
        const uint32_t o1 = __shfl_xor_sync(gmask, s1, o), o2 = __shfl_xor_sync(gmask, s2, o);
        const uint32_t lo = min(s1, o1), hi = max(s1, o1);
        s2 = min(hi, min(s2, o2));
        s1 = lo;
    }
    if (sub == 0)
    {
        ocb_top2 r;
        r.best_k = s1 == K4_NONE ? 0u : (s1 & K4_POS_MASK); // best_feat_idx starts at 0 (dense_stereo.cpp:253)
        r.best_d = s1 == K4_NONE ? (uint16_t)OCB_DIST_INF : (uint16_t)(s1 >> K4_POS_BITS);
        r.second_d = s2 == K4_NONE ? (uint16_t)OCB_DIST_INF : (uint16_t)(s2 >> K4_POS_BITS);
        out[list] = r;
    }
}

uint32_t k4_max_list_length()
{
    return K4_POS_MASK; // positions 0 .. 2^22 - 2; the all-ones key is "nothing yet"
}

int k4_launch(const void *d_q_rows, const void *d_c_rows, const uint32_t *d_list_query, const uint64_t *d_list_begin,
              const uint32_t *d_list_candidates, size_t n_lists, ocb_top2 *d_out, cudaStream_t stream)
{
    if (n_lists == 0)
        return 0;
    const unsigned grid = (unsigned)((n_lists + K4_LISTS_PER_CTA - 1) / K4_LISTS_PER_CTA);
    k4_lists_kernel<<<grid, K4_THREADS, 0, stream>>>(static_cast<const uint4 *>(d_q_rows),
                                                     static_cast<const uint4 *>(d_c_rows), d_list_query, d_list_begin,
                                                     d_list_candidates, (uint32_t)n_lists, d_out);
    count_launch();
    OCB_CUDA(cudaGetLastError());
    return 0;
}

} // namespace ocb
