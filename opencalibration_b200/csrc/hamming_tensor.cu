// K1T -- the Hamming top-2 search of one pair on the 5th-generation tensor cores (sm_100a: tcgen05.mma, TMEM, bulk TMA).
//
// Same contract as K1 (hamming_top2.cu; reference loop nest src/match/match_features.cpp:71-93): per query the first
// candidate position at minimum distance, that distance, the second-smallest distance with multiplicity; optionally,
// from the same sweep, the first query at minimum distance of every candidate (cross-check). Same records, bit for bit.
//
// Why a contraction is EXACT here. Map every descriptor bit b to s = 2 b - 1 in {-1, +1} (s8). Over the 512 positions
// of a padded row, equal bits contribute +1 and different bits -1 to the dot product, and the 26 padding bits are equal
// in every row, so    dot(q, c) = 512 - 2 * hamming(q, c)    exactly, in integers: tcgen05.mma.kind::i8 accumulates
// s8 x s8 products in s32 without rounding. north_star put this path on the integer pipes "because this is not a dense
// floating-point contraction"; measured on B200 the integer form peaks at 410 - 460 Gcmp/s, this
// one runs the 10k x 10k pair at > 1000 Gcmp/s with identical results, so it is the engine for large single pairs.
//
// Data flow
//   k1t_expand_kernel  64-byte rows -> s8 tiles of 128 rows in the K-major, no-swizzle canonical layout of a UMMA
//                      shared-memory descriptor: [K chunk of 16 B][row][16 B] (core matrix = 8 rows x 16 B contiguous;
//                      SBO = 128 B between 8-row groups, LBO = 2048 B between K chunks). A tile is 64 KB and arrives in
//                      shared memory with ONE bulk copy.
//   k1t_top2_kernel    CTA = (query tile of 128 rows, contiguous range of candidate tiles), 10 warps:
//                        warp 0    producer: bulk copies (cp.async.bulk + mbarrier) of the query tile, then of the
//                                  candidate tiles through a two-stage ring;
//                        warp 1    one thread issues, per candidate tile, 16 x tcgen05.mma (M 128 x N 128 x K 32) into
//                                  one of two 128-column TMEM accumulators and commits them to mbarriers (stage free,
//                                  accumulator full);
//                        warps 2-9 epilogue (two per TMEM lane quarter, one column half each): tcgen05.ld 32 columns
//                                  at a time (thread = query row), one IMAD turns the dot product into
//                                  v = distance << 20 | position, then K1's branch-free two-smallest update.
//                                  Cross-check: per candidate the warp maximum (REDUX) of (dot + 513) << 8 | (127 - row),
//                                  combined over the quarters through shared memory, one 64-bit atomic max per
//                                  candidate and CTA on the complemented (distance, query) key, as in K1.
//   k1t_finish_kernel  merges the candidate ranges of every query (positions are global, so packed values merge by
//                      min / max), writes the records and turns the column keys into query indices.
// Bound: L2 bandwidth at this tile shape (every CTA streams the candidate tiles: n1 / 128 x n2 x 512 B), then the
// epilogue's issue slots; the tensor pipe itself is about one third busy.
#include "ocb_internal.cuh"

#include <algorithm>
#include <type_traits>

namespace ocb
{
namespace
{
constexpr int T_ROWS = 128;                             // rows per tile (M and N of one MMA)
constexpr int T_KBYTES = 512;                           // s8 elements per row
constexpr uint32_t T_TILE_BYTES = T_ROWS * T_KBYTES;    // 64 KB
constexpr uint32_t T_LBO = T_ROWS * 16;                 // bytes between K chunks
constexpr uint32_t T_SBO = 8 * 16;                      // bytes between 8-row groups
constexpr int T_STAGES = 2;
constexpr int T_EPI_WARPS = 8;                         // epilogue warps: two per TMEM lane quarter (column halves)
constexpr int T_THREADS = (2 + T_EPI_WARPS) * 32;
constexpr uint32_t T_SHIFT = 20;
constexpr uint32_t T_NONE = 0xFFFFFFFFu;
constexpr uint32_t T_TMEM_COLS = 256;

// rows [n][8] u64 -> tiles [ceil(n/128)][32 K chunks][128 rows][16 B] of s8 (+1 / -1); rows past n: all -1
__global__ void __launch_bounds__(256)
    k1t_expand_kernel(const uint64_t *__restrict__ rows_a, uint32_t n_a, uint32_t n_a_padded, uint4 *__restrict__ tiles_a,
                      const uint64_t *__restrict__ rows_b, uint32_t n_b, uint32_t n_b_padded, uint4 *__restrict__ tiles_b)
{
    uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; // (row, K chunk) of set a, then of set b
    const uint64_t *rows = rows_a;
    uint32_t n = n_a;
    uint4 *tiles = tiles_a;
    if (g >= n_a_padded * 32u)
    {
        g -= n_a_padded * 32u;
        if (g >= n_b_padded * 32u)
            return;
        rows = rows_b, n = n_b, tiles = tiles_b;
    }
    const uint32_t kc = g & 31u, row = g >> 5;
    uint32_t bits16 = 0;
    if (row < n)
        bits16 = (uint32_t)(rows[(size_t)row * 8 + (kc >> 2)] >> ((kc & 3u) * 16u)) & 0xFFFFu;
    auto spread = [](uint32_t nib) { // 4 bits -> 4 bytes: 0x01 where the bit is set, 0xFF where it is clear
        const uint32_t b = (nib * 0x00204081u) & 0x01010101u;
        return 0xFFFFFFFFu - b * 0xFEu;
    };
    uint4 v;
    v.x = spread(bits16 & 15u), v.y = spread((bits16 >> 4) & 15u), v.z = spread((bits16 >> 8) & 15u),
    v.w = spread((bits16 >> 12) & 15u);
    const uint32_t tile = row / T_ROWS, r = row % T_ROWS;
    tiles[((size_t)tile * 32 + kc) * T_ROWS + r] = v;
}

// A barrier that does not flip within 4 M polls is a bug, not a delay: trap instead of hanging the device.
__device__ __forceinline__ void wait_or_trap(uint64_t *bar, uint32_t parity)
{
#pragma unroll 1
    for (uint32_t spin = 0; spin < (1u << 22); spin++)
        if (mbar_try_wait(bar, parity))
            return;
    asm volatile("trap;");
}

__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr)
{
    // cute::UMMA::SmemDescriptor: start address, leading / stride byte offsets (all >> 4), version 1, no swizzle
    uint64_t d = (uint64_t)((smem_addr >> 4) & 0x3FFFu);
    d |= (uint64_t)((T_LBO >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((T_SBO >> 4) & 0x3FFFu) << 32;
    d |= 1ull << 46;
    return d;
}

struct K1TParams
{
    const uint4 *q_tiles;       // expanded query tiles
    const uint4 *c_tiles;       // expanded candidate tiles
    uint32_t *part;             // [ranges][2 column halves][n1_padded][2] packed (s1, s2)
    unsigned long long *col64;  // [n2] complemented (distance, query) keys, zero before the launch; cross-check only
    uint32_t n1, n1_padded, n2, c_tiles_total, tiles_per_range;
};

template <bool COL> __global__ void __launch_bounds__(T_THREADS, 1) k1t_top2_kernel(const K1TParams P)
{
    extern __shared__ __align__(128) unsigned char k1t_smem[];
    unsigned char *sA = k1t_smem;
    unsigned char *sB = k1t_smem + T_TILE_BYTES;
    uint64_t *bars = reinterpret_cast<uint64_t *>(k1t_smem + (1 + T_STAGES) * T_TILE_BYTES);
    uint64_t *a_full = bars, *b_full = bars + 1, *b_empty = bars + 3, *d_full = bars + 5, *d_empty = bars + 7;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 9);
    uint32_t *colmin = reinterpret_cast<uint32_t *>(bars + 10); // [2][4][128]: cross-check, double-buffered per tile

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t qtile = blockIdx.x, range = blockIdx.y;
    const uint32_t t_begin = range * P.tiles_per_range;
    const uint32_t t_end = min(P.c_tiles_total, t_begin + P.tiles_per_range);
    const uint32_t ntiles = t_end > t_begin ? t_end - t_begin : 0;

    if (threadIdx.x == 0)
    {
        mbar_init(a_full, 1);
        for (int s = 0; s < T_STAGES; s++)
        {
            mbar_init(&b_full[s], 1);
            mbar_init(&b_empty[s], 1);
            mbar_init(&d_full[s], 1);
            mbar_init(&d_empty[s], T_EPI_WARPS);
        }
        mbar_fence_init();
    }
    if (warp == 1)
    {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(T_TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0)
    {
        if (lane == 0)
        {
            mbar_expect_tx(a_full, T_TILE_BYTES);
            bulk_g2s(sA, P.q_tiles + (size_t)qtile * (T_TILE_BYTES / 16), T_TILE_BYTES, a_full);
            for (uint32_t t = 0; t < ntiles; t++)
            {
                const uint32_t s = t % T_STAGES;
                if (t >= T_STAGES)
                    wait_or_trap(&b_empty[s], ((t / T_STAGES) - 1) & 1);
                mbar_expect_tx(&b_full[s], T_TILE_BYTES);
                bulk_g2s(sB + s * T_TILE_BYTES, P.c_tiles + (size_t)(t_begin + t) * (T_TILE_BYTES / 16), T_TILE_BYTES,
                         &b_full[s]);
            }
        }
    }
    else if (warp == 1)
    {
        if (lane == 0)
        {
            // cute::UMMA::InstrDescriptor: D = S32 (2 << 4), A and B signed 8 bit (1 << 7, 1 << 10), both K-major,
            // N >> 3 at bit 17, M >> 4 at bit 24
            const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(T_ROWS >> 3) << 17) |
                                   ((uint32_t)(T_ROWS >> 4) << 24);
            wait_or_trap(a_full, 0);
            const uint32_t a_addr = smem_u32(sA);
            for (uint32_t t = 0; t < ntiles; t++)
            {
                const uint32_t s = t % T_STAGES, buf = t & 1;
                wait_or_trap(&b_full[s], (t / T_STAGES) & 1);
                if (t >= 2)
                    wait_or_trap(&d_empty[buf], ((t >> 1) - 1) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t b_addr = smem_u32(sB + s * T_TILE_BYTES);
                const uint32_t d_tmem = tmem_base + buf * T_ROWS;
#pragma unroll
                for (uint32_t j = 0; j < T_KBYTES / 32; j++)
                {
                    const uint64_t adesc = umma_desc(a_addr + j * 2 * T_LBO), bdesc = umma_desc(b_addr + j * 2 * T_LBO);
                    const uint32_t accumulate = j > 0 ? 1u : 0u;
                    asm volatile("{\n\t"
                                 ".reg .pred p;\n\t"
                                 "setp.ne.b32 p, %4, 0;\n\t"
                                 "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
                                 "}\n" ::"r"(d_tmem),
                                 "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
                                 : "memory");
                }
                // both commits fire when the MMAs above have completed: the stage may be refilled, the accumulator read
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                                 smem_u32(&b_empty[s]))
                             : "memory");
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                                 smem_u32(&d_full[buf]))
                             : "memory");
            }
        }
    }
    else
    {
        // Eight epilogue warps: warp w reads the TMEM lanes of quarter w % 4 (a hardware rule) and the column half
        // (w - 2) / 4 of every accumulator, so that every scheduler of the SM has two warps to issue from.
        const uint32_t quarter = warp & 3u;
        const uint32_t half = (warp - 2u) >> 2;
        const uint32_t row = quarter * 32u + lane; // query row of this thread within the tile
        const uint32_t qpos = qtile * T_ROWS + row;
        // cross-check key of (this query, a candidate) = (dot + 513) << 8 | (127 - row): its MAXIMUM over the queries is
        // the smallest distance and, among equals, the smallest row; 0 = nothing (padding rows contribute 0)
        const uint32_t key_mul = qpos < P.n1 ? 256u : 0u;
        const uint32_t key_add = qpos < P.n1 ? ((513u << 8) | (127u - row)) : 0u;
        uint32_t s1 = T_NONE, s2 = T_NONE;
        auto sweep = [&](uint32_t t, auto partial) {
            const uint32_t buf = t & 1;
            const uint32_t pos0 = (t_begin + t) * T_ROWS;
#pragma unroll 1
            for (uint32_t c0 = half * 64u; c0 < half * 64u + 64u; c0 += 32)
            {
                uint32_t d[32];
                const uint32_t taddr = tmem_base + ((quarter * 32u) << 16) + buf * T_ROWS + c0;
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                             "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                             "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
                             : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6]), "=r"(d[7]),
                               "=r"(d[8]), "=r"(d[9]), "=r"(d[10]), "=r"(d[11]), "=r"(d[12]), "=r"(d[13]), "=r"(d[14]),
                               "=r"(d[15]), "=r"(d[16]), "=r"(d[17]), "=r"(d[18]), "=r"(d[19]), "=r"(d[20]), "=r"(d[21]),
                               "=r"(d[22]), "=r"(d[23]), "=r"(d[24]), "=r"(d[25]), "=r"(d[26]), "=r"(d[27]), "=r"(d[28]),
                               "=r"(d[29]), "=r"(d[30]), "=r"(d[31])
                             : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                // v = (512 - dot) / 2 << 20 | position; dot is even, so (512 - dot) << 19 is exact: one IMAD
                const uint32_t base = (512u << (T_SHIFT - 1)) + pos0 + c0;
                uint32_t *cm = colmin + (t & 1u) * 512u + quarter * 128u + c0;
#pragma unroll
                for (int i = 0; i < 32; i++)
                {
                    uint32_t v = d[i] * (0u - (1u << (T_SHIFT - 1))) + (base + (uint32_t)i);
                    if constexpr (decltype(partial)::value) // only the last candidate tile can hold padding columns
                        v = pos0 + c0 + (uint32_t)i < P.n2 ? v : T_NONE;
                    s2 = min(s2, max(s1, v));
                    s1 = min(s1, v);
                    if constexpr (COL)
                    {
                        const uint32_t m = __reduce_max_sync(0xFFFFFFFFu, d[i] * key_mul + key_add);
                        if (lane == 0)
                            cm[i] = m; // the warp's best (distance, row) for candidate pos0 + c0 + i
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0)
                mbar_arrive(&d_empty[buf]);
            if constexpr (COL)
            {
                // the eight epilogue warps meet (named barrier 1); thread k < 128 then owns candidate pos0 + k. Buffer
                // t & 1 is written again in tile t + 2, i.e. after the barrier of tile t + 1, which that thread reaches
                // only after this flush.
                asm volatile("bar.sync 1, 256;" ::: "memory");
                const uint32_t k = (warp - 2u) * 32u + lane;
                if (k < T_ROWS)
                {
                    const uint32_t *all = colmin + (t & 1u) * 512u;
                    const uint32_t m = max(max(all[k], all[128 + k]), max(all[256 + k], all[384 + k]));
                    if (pos0 + k < P.n2 && m != 0u)
                    {
                        const unsigned long long key = ((unsigned long long)((1025u - (m >> 8)) >> 1) << 32) |
                                                       (unsigned long long)(qtile * T_ROWS + 127u - (m & 0xFFu));
                        red_max_u64_global(&P.col64[pos0 + k], ~key);
                    }
                }
            }
        };
        for (uint32_t t = 0; t < ntiles; t++)
        {
            wait_or_trap(&d_full[t & 1], (t >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if ((t_begin + t + 1) * T_ROWS <= P.n2)
                sweep(t, std::false_type());
            else
                sweep(t, std::true_type());
        }
        uint32_t *out = P.part + ((size_t)(range * 2u + half) * P.n1_padded + qpos) * 2;
        out[0] = s1, out[1] = s2;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(T_TMEM_COLS));
}

__global__ void __launch_bounds__(256)
    k1t_finish_kernel(const uint32_t *__restrict__ part, uint32_t ranges, uint32_t n1_padded, uint32_t n1,
                      ocb_top2 *__restrict__ out, const unsigned long long *__restrict__ col64, uint32_t n2,
                      uint32_t *__restrict__ col_out)
{
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g < n1)
    {
        uint32_t s1 = T_NONE, s2 = T_NONE;
        for (uint32_t s = 0; s < ranges; s++)
        {
            const uint32_t *p = part + ((size_t)s * n1_padded + g) * 2;
            const uint32_t a = p[0], b = p[1];
            s2 = min(s2, max(s1, a));
            s1 = min(s1, a);
            s2 = min(s2, max(s1, b));
            s1 = min(s1, b);
        }
        ocb_top2 r;
        r.best_k = s1 == T_NONE ? 0u : (s1 & ((1u << T_SHIFT) - 1u)); // feature_match best_match{i, 0, inf} (:74)
        r.best_d = s1 == T_NONE ? (uint16_t)OCB_DIST_INF : (uint16_t)(s1 >> T_SHIFT);
        r.second_d = s2 == T_NONE ? (uint16_t)OCB_DIST_INF : (uint16_t)(s2 >> T_SHIFT);
        out[g] = r;
    }
    if (col_out && g < n2)
        col_out[g] = (uint32_t)(~col64[g]); // zero (nothing seen) -> OCB_NO_INDEX
}

struct K1TLayout
{
    uint32_t q_tiles, c_tiles, n1p, n2p, ranges, per;
    size_t o_qt, o_ct, o_part, o_col, total;
};

K1TLayout k1t_layout(size_t n1, size_t n2, bool col, int sms)
{
    K1TLayout L;
    L.q_tiles = (uint32_t)((n1 + T_ROWS - 1) / T_ROWS), L.c_tiles = (uint32_t)((n2 + T_ROWS - 1) / T_ROWS);
    L.n1p = L.q_tiles * T_ROWS, L.n2p = L.c_tiles * T_ROWS;
    // candidate ranges: whole waves of one CTA per SM, as few tile-loads of the query tile as that allows
    uint32_t best_ranges = 1;
    double best = 1e30;
    for (uint32_t s = 1; s <= std::min<uint32_t>(std::max<uint32_t>(L.c_tiles, 1), 64); s++)
    {
        const uint32_t per = (L.c_tiles + s - 1) / s, ranges = (L.c_tiles + per - 1) / per;
        const double waves = (double)(((uint64_t)L.q_tiles * ranges + sms - 1) / sms);
        const double cost = waves * (per + 1.5); // tiles per CTA + the query tile load and the pipeline fill
        if (cost < best)
            best = cost, best_ranges = ranges;
    }
    L.per = (L.c_tiles + best_ranges - 1) / best_ranges;
    L.ranges = (L.c_tiles + L.per - 1) / L.per;
    size_t off = 0;
    auto take = [&off](size_t bytes) {
        const size_t o = off;
        off = (off + bytes + 255) / 256 * 256;
        return o;
    };
    L.o_qt = take((size_t)L.q_tiles * T_TILE_BYTES);
    L.o_ct = take((size_t)L.c_tiles * T_TILE_BYTES);
    L.o_part = take((size_t)L.ranges * 2 * L.n1p * 2 * sizeof(uint32_t));
    L.o_col = take(col ? n2 * sizeof(unsigned long long) : 0);
    L.total = off;
    return L;
}
} // namespace

bool k1t_supports(size_t n1, size_t n2)
{
    return n1 >= 1 && n2 >= 1 && n2 < (1u << T_SHIFT) && n1 < (1u << 24);
}

size_t k1t_workspace_bytes(size_t n1, size_t n2, bool col)
{
    if (!k1t_supports(n1, n2))
        return 0;
    // the number of candidate ranges depends on the SM count: take the worst case over the counts the planner may see
    size_t worst = 0;
    for (int sms : {148, 132, 108, 64, 1})
        worst = std::max(worst, k1t_layout(n1, n2, col, sms).total);
    return worst + 256;
}

int k1t_launch(const void *d_q, size_t n1, const void *d_c, size_t n2, ocb_top2 *d_out, uint32_t *d_col_best_q,
               void *d_workspace, int sms, cudaStream_t stream)
{
    const bool col = d_col_best_q != nullptr;
    const K1TLayout L = k1t_layout(n1, n2, col, sms);
    char *ws = static_cast<char *>(d_workspace);
    uint4 *qt = reinterpret_cast<uint4 *>(ws + L.o_qt), *ct = reinterpret_cast<uint4 *>(ws + L.o_ct);
    K1TParams P;
    P.q_tiles = qt, P.c_tiles = ct;
    P.part = reinterpret_cast<uint32_t *>(ws + L.o_part);
    P.col64 = col ? reinterpret_cast<unsigned long long *>(ws + L.o_col) : nullptr;
    P.n1 = (uint32_t)n1, P.n1_padded = L.n1p, P.n2 = (uint32_t)n2, P.c_tiles_total = L.c_tiles, P.tiles_per_range = L.per;
    if (col)
        OCB_CUDA(cudaMemsetAsync(ws + L.o_col, 0, n2 * sizeof(unsigned long long), stream));
    const uint32_t expand_threads = (L.n1p + L.n2p) * 32u;
    k1t_expand_kernel<<<(expand_threads + 255) / 256, 256, 0, stream>>>(static_cast<const uint64_t *>(d_q), (uint32_t)n1,
                                                                       L.n1p, qt, static_cast<const uint64_t *>(d_c),
                                                                       (uint32_t)n2, L.n2p, ct);
    const size_t smem = (size_t)(1 + T_STAGES) * T_TILE_BYTES + 128 + (col ? 2 * 4 * 128 * sizeof(uint32_t) : 0);
    if (col)
    {
        OCB_CUDA(cudaFuncSetAttribute(k1t_top2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k1t_top2_kernel<true><<<dim3(L.q_tiles, L.ranges), T_THREADS, smem, stream>>>(P);
    }
    else
    {
        OCB_CUDA(cudaFuncSetAttribute(k1t_top2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k1t_top2_kernel<false><<<dim3(L.q_tiles, L.ranges), T_THREADS, smem, stream>>>(P);
    }
    const uint32_t finish = (uint32_t)std::max(n1, col ? n2 : (size_t)0);
    k1t_finish_kernel<<<(finish + 255) / 256, 256, 0, stream>>>(P.part, L.ranges * 2, L.n1p, (uint32_t)n1, d_out, P.col64,
                                                              (uint32_t)n2, d_col_best_q);
    count_launch(3);
    OCB_CUDA(cudaGetLastError());
    return 0;
}

} // namespace ocb
