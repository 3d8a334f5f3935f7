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
// Data flow (three launches per call: expansion, search, finish)
//   expansion   64-byte candidate rows -> s8 tiles in the K-major, no-swizzle canonical layout of a UMMA shared-memory
//               descriptor: [K chunk of 16 B][row][16 B] (core matrix = 8 rows x 16 B contiguous; SBO = 128 B between
//               8-row groups, LBO = rows x 16 B between K chunks); a tile arrives in shared memory with ONE bulk copy.
//   search      persistent CTAs, warp-specialised: a bulk-copy producer (cp.async.bulk + mbarrier), one MMA-issuing warp
//               (tcgen05.mma.kind::i8 into TMEM accumulators, tcgen05.commit to mbarriers), 8 or 16 epilogue warps
//               (tcgen05.ld, thread = query row: packed keys, branch-free two-smallest), a cross-check flush warp.
//   finish      merges the partial pairs of the CTAs that worked on a query row, writes the records and turns the
//               column keys into query indices.
// Four forms of the search kernel are kept (option k1t_variant; every test of tests/test_gpu_match.py runs on each):
//   1  first form    CTA = (128-query tile, candidate range), A and B from shared memory, 8 epilogue warps.
//   2  second form   16 epilogue warps, persistent spans over the linearised (query tile, candidate step) space, MMAs
//                    issued by an ELECTED lane of a warp-uniform loop (under `lane == 0` ptxas wraps every tcgen05.mma in
//                    an ELECT / R2UR.BROADCAST / BRA.U.ANY waterfall: ~110 cycles of issue per MMA), 16-bit cross-check
//                    keys handed to a flush warp through mbarriers instead of a CTA-wide barrier per step.
//   3                the same with two query tiles per CTA against steps of 64 candidates (half the L2 stream; slower:
//                    an A read from shared memory costs 32 wavefronts per MMA whatever N is).
//   4  third form    DEFAULT: the query operand lives in TENSOR MEMORY (tcgen05.mma with A from TMEM): the epilogue
//                    warps expand their own query rows in registers and tcgen05.st them, two query tiles per CTA, the
//                    227 KB of shared memory are a six-stage ring of candidate steps, top-2 on packed 16-bit keys
//                    (VIMNMX.U16x2 / VIMNMX3.U16x2).
// What bounds it (ocb_probe_umma, include/ocb_probe.h): an accumulator accepts a dependent MMA
// only every ~150 cycles whatever N is, and 256 of the 512 TMEM columns hold A, so two accumulators of 64 columns are in
// flight per buffer: an M 128 x N 64 x K 32 MMA retires every ~64 cycles against 32 of arithmetic -- half the tensor pipe.
// The full rate needs N = 256 per MMA, i.e. a CTA pair (cta_group::2); not built.
#include "ocb_internal.cuh"

#include <algorithm>
#include <type_traits>
#include <vector>

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
constexpr int K1T_DEFAULT_VARIANT = 4; // form of the search kernel when the option k1t_variant is 0 (see k1t_launch)

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

// elect.sync: true in exactly one lane of the (converged) warp; unlike `lane == 0` the compiler knows it
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred = 0;
    asm volatile("{\n\t"
                 ".reg .pred p;\n\t"
                 "elect.sync _|p, 0xFFFFFFFF;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t"
                 "}\n"
                 : "=r"(pred));
    return pred != 0;
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

// ---------------------------------------------------------------------------------------------------------------------
// K1T second form (round 2, second session): the same contraction with
//   * 16 epilogue warps (four per scheduler instead of two: the first form's epilogue issued on 57 % of the cycles with
//     two warps per scheduler and bounded the step at ~2700 cycles against 1024 of the MMAs),
//   * block-local signed keys  w = dot * -2^19 + column  (one IMAD with an immediate addend, no per-value add for the
//     position; the two smallest of a 32-column block are rebased once per block),
//   * the cross-check handed to a separate warp through 16-bit keys and two mbarrier pairs instead of a CTA-wide named
//     barrier per candidate step (the epilogue warps no longer run in lock-step),
//   * the accumulator released right after tcgen05.ld (before the arithmetic),
//   * MT query tiles per CTA against candidate steps of NT rows (MT x NT = 128 accumulator columns per buffer):
//     <1, 128> is the first form's shape, <2, 64> halves the bytes streamed from L2 per comparison,
//   * persistent CTAs over the linearised (query tile, candidate step) space: every SM gets the same number of steps,
//     whatever the shape; a CTA whose span crosses into the next query tile reloads A once.
// Roles: warps 0-15 epilogue (TMEM lane quarter = warp % 4, 32-column block = warp / 4), warp 16 bulk-copy producer,
// warp 17 MMA issuer, warp 18 cross-check flush.
constexpr int V_EPI_WARPS = 16;
constexpr int V_THREADS = (V_EPI_WARPS + 3) * 32;
constexpr uint32_t V_A_TILE_BYTES = 128 * T_KBYTES; // one query tile (M = 128)
constexpr uint32_t V_BAR_BYTES = 256;
constexpr uint32_t V_COLMIN_BYTES = 2 * 1024;        // [2 buffers][MT * 4 row groups][NT] u16

struct K1T2Params
{
    const uint4 *q_tiles;      // expanded query tiles of 128 rows
    const uint4 *c_tiles;      // expanded candidate tiles of NT rows
    uint32_t *part;            // [slot * BPM + block][n1_padded][2] packed (s1, s2)
    unsigned long long *col64; // [n2] complemented (distance, query) keys, zeroed by the expansion kernel
    uint32_t n1, n1_padded, n2;
    uint32_t c_steps;          // candidate steps per query tile group = ceil(n2 / NT)
    uint32_t total_steps;      // query tile groups x c_steps
    uint32_t per;              // linear steps per CTA
};

template <int NT> __device__ __forceinline__ uint64_t umma_desc_lbo(uint32_t smem_addr)
{
    uint64_t d = (uint64_t)((smem_addr >> 4) & 0x3FFFu);
    d |= (uint64_t)(((uint32_t)NT * 16u >> 4) & 0x3FFFu) << 16; // bytes between K chunks of a tile of NT rows
    d |= (uint64_t)((T_SBO >> 4) & 0x3FFFu) << 32;
    d |= 1ull << 46;
    return d;
}

// rows [n][8] u64 -> tiles [ceil(n / TR)][32 K chunks][TR rows][16 B] of s8; rows past n: all -1. Thread = (row, quarter
// of the K chunks); consecutive lanes write consecutive 16-byte pieces. Also zeroes the cross-check keys.
__global__ void __launch_bounds__(128)
    k1t2_expand_kernel(const uint4 *__restrict__ rows_a, uint32_t n_a, uint32_t tiles_a, uint4 *__restrict__ out_a,
                       const uint4 *__restrict__ rows_b, uint32_t n_b, uint32_t rows_per_tile_b, uint32_t rows_b_padded,
                       uint4 *__restrict__ out_b, unsigned long long *__restrict__ col64, uint32_t n_col)
{
    // blockIdx.x: 128-row slab of set a (tiles_a of them), then of set b; blockIdx.y: quarter of the 32 K chunks
    const uint32_t slab = blockIdx.x, quarter = blockIdx.y;
    const bool is_a = slab < tiles_a;
    const uint32_t row = (is_a ? slab : slab - tiles_a) * 128u + threadIdx.x;
    const uint4 *rows = is_a ? rows_a : rows_b;
    const uint32_t n = is_a ? n_a : n_b;
    const uint32_t tr = is_a ? 128u : rows_per_tile_b;
    uint4 *out = is_a ? out_a : out_b;
    if (!is_a && quarter == 0 && col64 != nullptr)
    {
        if (row < n_col)
            col64[row] = 0ull;
    }
    if (!is_a && row >= rows_b_padded) // the last 128-row slab of set b may reach past its last tile (tiles of 64 rows)
        return;
    uint4 bits = make_uint4(0u, 0u, 0u, 0u);
    if (row < n)
        bits = rows[(size_t)row * 4 + quarter]; // 128 bits = 8 K chunks of 16 bits
    auto spread = [](uint32_t nib) { // 4 bits -> 4 bytes: 0x01 where the bit is set, 0xFF where it is clear
        const uint32_t b = (nib * 0x00204081u) & 0x01010101u;
        return 0xFFFFFFFFu - b * 0xFEu;
    };
    const uint32_t tile = row / tr, r = row % tr;
    uint4 *dst = out + ((size_t)tile * 32 + quarter * 8u) * tr + r;
    const uint32_t w[4] = {bits.x, bits.y, bits.z, bits.w};
#pragma unroll
    for (int k = 0; k < 8; k++)
    {
        const uint32_t bits16 = (w[k >> 1] >> ((k & 1) * 16)) & 0xFFFFu;
        uint4 v;
        v.x = spread(bits16 & 15u), v.y = spread((bits16 >> 4) & 15u), v.z = spread((bits16 >> 8) & 15u),
        v.w = spread((bits16 >> 12) & 15u);
        dst[(size_t)k * tr] = v;
    }
}

template <bool COL, int MT, int NT, int STAGES>
__global__ void __launch_bounds__(V_THREADS, 1) k1t2_top2_kernel(const K1T2Params P)
{
    static_assert(MT * NT == 128, "one accumulator buffer = 128 TMEM columns");
    constexpr uint32_t A_BYTES = (uint32_t)MT * V_A_TILE_BYTES;
    constexpr uint32_t B_BYTES = (uint32_t)NT * T_KBYTES;
    constexpr uint32_t BPM = NT / 32;  // 32-column blocks per query tile
    constexpr uint32_t G = MT * 4;     // row groups of 32 query rows
    constexpr uint32_t QROWS = MT * 128;
    extern __shared__ __align__(128) unsigned char k1t2_smem[];
    unsigned char *sA = k1t2_smem;
    unsigned char *sB = k1t2_smem + A_BYTES;
    uint64_t *bars = reinterpret_cast<uint64_t *>(k1t2_smem + A_BYTES + STAGES * B_BYTES);
    uint64_t *a_full = bars, *a_empty = bars + 1, *d_full = bars + 2, *d_empty = bars + 4, *c_full = bars + 6,
             *c_empty = bars + 8, *b_full = bars + 10, *b_empty = bars + 10 + STAGES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 10 + 2 * STAGES);
    static_assert((10 + 2 * STAGES + 1) * 8 <= V_BAR_BYTES, "barrier block");
    uint16_t *colmin = reinterpret_cast<uint16_t *>(k1t2_smem + A_BYTES + STAGES * B_BYTES + V_BAR_BYTES);

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t l_begin = blockIdx.x * P.per;
    const uint32_t l_end = min(P.total_steps, l_begin + P.per);

    if (threadIdx.x == 0)
    {
        mbar_init(a_full, 1);
        mbar_init(a_empty, 1);
        for (int s = 0; s < 2; s++)
        {
            mbar_init(&d_full[s], 1);
            mbar_init(&d_empty[s], V_EPI_WARPS);
            mbar_init(&c_full[s], V_EPI_WARPS);
            mbar_init(&c_empty[s], 1);
        }
        for (int s = 0; s < STAGES; s++)
        {
            mbar_init(&b_full[s], 1);
            mbar_init(&b_empty[s], 1);
        }
        mbar_fence_init();
    }
    if (warp == V_EPI_WARPS + 1)
    {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(T_TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == V_EPI_WARPS)
    {
        // ---- producer: A once per segment (a segment = the part of this CTA's span inside one query tile group), B
        // steps through the ring. The A load of a later segment is issued as late as the ring allows, so that the
        // candidate steps of the new segment are already on their way while the last MMAs of the old one drain.
        if (lane == 0)
        {
            uint32_t t = 0, seg = 0;
            for (uint32_t l = l_begin; l < l_end; seg++)
            {
                const uint32_t qb = l / P.c_steps, ct = l - qb * P.c_steps;
                const uint32_t n = min(P.c_steps - ct, l_end - l);
                const uint32_t a_at = seg == 0 ? 0u : min((uint32_t)STAGES - 1u, n - 1u);
                for (uint32_t i = 0; i < n; i++, t++)
                {
                    if (i == a_at)
                    {
                        if (seg > 0)
                            wait_or_trap(a_empty, (seg - 1) & 1);
                        mbar_expect_tx(a_full, A_BYTES);
                        bulk_g2s(sA, P.q_tiles + (size_t)qb * (A_BYTES / 16), A_BYTES, a_full);
                    }
                    const uint32_t s = t % STAGES;
                    if (t >= (uint32_t)STAGES)
                        wait_or_trap(&b_empty[s], ((t / STAGES) - 1) & 1);
                    mbar_expect_tx(&b_full[s], B_BYTES);
                    bulk_g2s(sB + s * B_BYTES, P.c_tiles + (size_t)(ct + i) * (B_BYTES / 16), B_BYTES, &b_full[s]);
                }
                l += n;
            }
        }
    }
    else if (warp == V_EPI_WARPS + 1)
    {
        // ---- MMA issuer. The WHOLE warp runs the loop (waits included) and one elected lane issues: with the loop under
        // `lane == 0` ptxas treats every descriptor as divergent and wraps each tcgen05.mma in an ELECT / R2UR.BROADCAST /
        // BRA.U.ANY waterfall -- ~12 dependent instructions, ~110 cycles per MMA, which bounded the first form (ncu:
        // the issuing warp never waits on a barrier, the tensor pipe is 27 - 40 % busy). Warp-uniform control flow keeps
        // the descriptors in uniform registers.
        // cute::UMMA::InstrDescriptor: D = S32 (2 << 4), A and B signed 8 bit (1 << 7, 1 << 10), both K-major,
        // N >> 3 at bit 17, M >> 4 at bit 24
        const uint32_t idesc =
            (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t a_addr = smem_u32(sA);
        const uint32_t b_addr0 = smem_u32(sB);
        uint32_t t = 0, seg = 0;
        for (uint32_t l = l_begin; l < l_end; seg++)
        {
            const uint32_t qb = l / P.c_steps, ct = l - qb * P.c_steps;
            const uint32_t n = min(P.c_steps - ct, l_end - l);
            wait_or_trap(a_full, seg & 1);
            for (uint32_t i = 0; i < n; i++, t++)
            {
                const uint32_t s = t % STAGES, buf = t & 1;
                wait_or_trap(&b_full[s], (t / STAGES) & 1);
                if (t >= 2)
                    wait_or_trap(&d_empty[buf], ((t >> 1) - 1) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (elect_one())
                {
                    const uint32_t b_addr = b_addr0 + s * B_BYTES;
                    // K steps outermost: consecutive MMAs go to different accumulators (the MT query tiles)
#pragma unroll
                    for (uint32_t j = 0; j < T_KBYTES / 32; j++)
                    {
#pragma unroll
                        for (uint32_t m = 0; m < (uint32_t)MT; m++)
                        {
                            const uint32_t d_tmem = tmem_base + buf * 128u + m * (uint32_t)NT;
                            const uint64_t adesc = umma_desc_lbo<128>(a_addr + m * V_A_TILE_BYTES + j * 2 * (128u * 16u));
                            const uint64_t bdesc = umma_desc_lbo<NT>(b_addr + j * 2 * ((uint32_t)NT * 16u));
                            const uint32_t accumulate = j > 0 ? 1u : 0u;
                            asm volatile("{\n\t"
                                         ".reg .pred p;\n\t"
                                         "setp.ne.b32 p, %4, 0;\n\t"
                                         "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
                                         "}\n" ::"r"(d_tmem),
                                         "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
                                         : "memory");
                        }
                    }
                    // both commits fire when the MMAs above have completed: the stage may be refilled, the accumulator read
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                                     smem_u32(&b_empty[s]))
                                 : "memory");
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                                     smem_u32(&d_full[buf]))
                                 : "memory");
                    // A may be overwritten once every MMA of this segment has completed
                    if (i + 1 == n)
                        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                                         smem_u32(a_empty))
                                     : "memory");
                }
                __syncwarp();
            }
            l += n;
        }
    }
    else if (warp == V_EPI_WARPS + 2)
    {
        // ---- cross-check flush: per candidate of a step, the best (distance, first query row) over the row groups, then
        // ONE 64-bit atomic max per candidate and CTA on the complemented (distance, query) key (K1's global protocol)
        if constexpr (COL)
        {
            uint32_t t = 0;
            for (uint32_t l = l_begin; l < l_end;)
            {
                const uint32_t qb = l / P.c_steps, ct = l - qb * P.c_steps;
                const uint32_t n = min(P.c_steps - ct, l_end - l);
                for (uint32_t i = 0; i < n; i++, t++)
                {
                    const uint32_t b = t & 1;
                    wait_or_trap(&c_full[b], (t >> 1) & 1);
                    const uint16_t *cm = colmin + b * (G * NT);
                    uint32_t best[NT / 32];
#pragma unroll
                    for (uint32_t k = 0; k < (uint32_t)NT / 32; k++)
                    {
                        const uint32_t c = k * 32u + lane;
                        uint32_t m = 0;
#pragma unroll
                        for (uint32_t g = 0; g < G; g++)
                        {
                            const uint32_t key = cm[g * NT + c];
                            // (dot / 2 + 257) << 8 | (255 - row within the query tile group); 0 = nothing seen
                            const uint32_t cand = ((key >> 5) << 8) | (255u - (g * 32u + 31u - (key & 31u)));
                            m = max(m, key ? cand : 0u);
                        }
                        best[k] = m;
                    }
                    __syncwarp();
                    if (lane == 0)
                        mbar_arrive(&c_empty[b]);
#pragma unroll
                    for (uint32_t k = 0; k < (uint32_t)NT / 32; k++)
                    {
                        const uint32_t pos = (ct + i) * (uint32_t)NT + k * 32u + lane;
                        if (best[k] != 0u && pos < P.n2)
                        {
                            const unsigned long long key =
                                ((unsigned long long)(513u - (best[k] >> 8)) << 32) |
                                (unsigned long long)(qb * QROWS + 255u - (best[k] & 255u));
                            red_max_u64_global(&P.col64[pos], ~key);
                        }
                    }
                }
                l += n;
            }
        }
    }
    else
    {
        // ---- epilogue: warp w reads the TMEM lanes of quarter w % 4 (a hardware rule) and the 32-column block w / 4 of
        // every accumulator buffer: block = (query tile m of the group, column block of the step)
        const uint32_t quarter = warp & 3u, blk = warp >> 2;
        const uint32_t mt = blk / BPM, cblk = blk % BPM;
        const uint32_t row_local = mt * 128u + quarter * 32u + lane;
        const uint32_t g = mt * 4u + quarter;
        uint32_t t = 0;
        for (uint32_t l = l_begin; l < l_end;)
        {
            const uint32_t qb = l / P.c_steps, ct = l - qb * P.c_steps;
            const uint32_t n = min(P.c_steps - ct, l_end - l);
            const uint32_t qpos = qb * QROWS + row_local;
            // cross-check key of (this query, a candidate) within the warp = (dot / 2 + 257) << 5 | (31 - lane): its
            // MAXIMUM over the lanes is the smallest distance and, among equals, the smallest row; 0 = nothing (padding
            // rows contribute 0). dot is even, so dot * 16 + 8224 is exact.
            const uint32_t key_mul = qpos < P.n1 ? 16u : 0u;
            const uint32_t key_add = qpos < P.n1 ? (8224u + 31u - lane) : 0u;
            uint32_t s1 = T_NONE, s2 = T_NONE;
            auto step = [&](uint32_t cstep, auto partial) {
                const uint32_t buf = t & 1;
                const uint32_t pos0 = cstep * (uint32_t)NT + cblk * 32u; // position of this block's first column
                wait_or_trap(&d_full[buf], (t >> 1) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                uint32_t d[32];
                const uint32_t taddr = tmem_base + ((quarter * 32u) << 16) + buf * 128u + blk * 32u;
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
                // the values are in registers: the accumulator buffer may be overwritten by the MMAs of step t + 2
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0)
                    mbar_arrive(&d_empty[buf]);
                uint16_t *cm = colmin + (t & 1u) * (G * NT) + g * NT + cblk * 32u;
                if constexpr (COL)
                    if (t >= 2)
                        wait_or_trap(&c_empty[t & 1], ((t >> 1) - 1) & 1);
                // block-local signed key  w = dot * -2^19 + i = (distance - 256) << 20 | i : ordered like (distance, column)
                int32_t b1 = 0x7FFFFFFF, b2 = 0x7FFFFFFF;
                auto key_of = [&](int i) {
                    int32_t w = (int32_t)d[i] * (-(1 << (T_SHIFT - 1))) + i;
                    if constexpr (decltype(partial)::value) // only the last candidate step can hold padding columns
                        w = pos0 + (uint32_t)i < P.n2 ? w : 0x7FFFFFFF;
                    return w;
                };
#pragma unroll
                for (int i = 0; i < 32; i += 4)
                {
                    // two values at a time: with lo <= hi the two smallest of {b1, b2, lo, hi} are min(b1, lo) and
                    // min3(max(b1, lo), b2, hi): five min / max operations per two values (VIMNMX, VIMNMX3)
#pragma unroll
                    for (int k = 0; k < 4; k += 2)
                    {
                        const int32_t wa = key_of(i + k), wb = key_of(i + k + 1);
                        const int32_t lo = min(wa, wb), hi = max(wa, wb);
                        b2 = __vimin3_s32(max(b1, lo), b2, hi);
                        b1 = min(b1, lo);
                    }
                    if constexpr (COL)
                    {
                        // the warp's best (distance, row) for the candidates pos0 + i .. i + 3: four 16-bit keys, one store
                        uint32_t m[4];
#pragma unroll
                        for (int k = 0; k < 4; k++)
                            m[k] = __reduce_max_sync(0xFFFFFFFFu, d[i + k] * key_mul + key_add);
                        if (lane == 0)
                            *reinterpret_cast<uint2 *>(cm + i) = make_uint2(m[0] | (m[1] << 16), m[2] | (m[3] << 16));
                    }
                }
                if constexpr (COL)
                    if (lane == 0)
                        mbar_arrive(&c_full[t & 1]); // release: the flush warp sees the 32 stores above
                // rebase the block's two smallest to v = distance << 20 | position and merge them into the running pair
                const uint32_t off = (1u << 28) + pos0;
                uint32_t g1 = (uint32_t)b1 + off, g2 = (uint32_t)b2 + off;
                if constexpr (decltype(partial)::value)
                {
                    g1 = b1 == 0x7FFFFFFF ? T_NONE : g1;
                    g2 = b2 == 0x7FFFFFFF ? T_NONE : g2;
                }
                s2 = min(s2, max(s1, g1));
                s1 = min(s1, g1);
                s2 = min(s2, max(s1, g2));
                s1 = min(s1, g2);
            };
            for (uint32_t i = 0; i < n; i++, t++)
            {
                if ((ct + i + 1) * (uint32_t)NT <= P.n2)
                    step(ct + i, std::false_type());
                else
                    step(ct + i, std::true_type());
            }
            // this CTA is the slot-th one working on query tile group qb
            const uint32_t slot = blockIdx.x - (qb * P.c_steps) / P.per;
            uint32_t *out = P.part + ((size_t)(slot * BPM + cblk) * P.n1_padded + qpos) * 2;
            out[0] = s1, out[1] = s2;
            l += n;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == V_EPI_WARPS + 1)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(T_TMEM_COLS));
}

__global__ void __launch_bounds__(256)
    k1t2_finish_kernel(const uint32_t *__restrict__ part, uint32_t n1_padded, uint32_t n1, uint32_t qrows, uint32_t bpm,
                       uint32_t c_steps, uint32_t per, ocb_top2 *__restrict__ out,
                       const unsigned long long *__restrict__ col64, uint32_t n2, uint32_t *__restrict__ col_out)
{
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long column_key = 0ull; // asked for first: its round trip overlaps the ones of the partial pairs
    if (col_out && g < n2)
        asm volatile("ld.global.u64 %0, [%1];" : "=l"(column_key) : "l"(col64 + g));
    if (g < n1)
    {
        // the CTAs whose spans touch this row's query tile group each left bpm partial pairs
        const uint32_t qb = g / qrows;
        const uint32_t first = (qb * c_steps) / per, last = ((qb + 1) * c_steps - 1) / per;
        const uint32_t slots = (last - first + 1) * bpm;
        uint32_t s1 = T_NONE, s2 = T_NONE;
        // The kernel is pure latency: eight loads in flight per thread. (Written as volatile loads from clamped, always
        // valid addresses: left to itself ptxas issues one load, consumes it, issues the next -- ten L2 round trips.)
        for (uint32_t s0 = 0; s0 < slots; s0 += 8)
        {
            uint2 p[8];
#pragma unroll
            for (uint32_t k = 0; k < 8; k++)
            {
                const uint32_t *src = part + ((size_t)min(s0 + k, slots - 1u) * n1_padded + g) * 2;
                asm volatile("ld.global.v2.u32 {%0, %1}, [%2];" : "=r"(p[k].x), "=r"(p[k].y) : "l"(src));
            }
#pragma unroll
            for (uint32_t k = 0; k < 8; k++)
            {
                const uint32_t a = s0 + k < slots ? p[k].x : T_NONE, b = s0 + k < slots ? p[k].y : T_NONE;
                s2 = min(s2, max(s1, a));
                s1 = min(s1, a);
                s2 = min(s2, max(s1, b));
                s1 = min(s1, b);
            }
        }
        ocb_top2 r;
        r.best_k = s1 == T_NONE ? 0u : (s1 & ((1u << T_SHIFT) - 1u)); // feature_match best_match{i, 0, inf} (:74)
        r.best_d = s1 == T_NONE ? (uint16_t)OCB_DIST_INF : (uint16_t)(s1 >> T_SHIFT);
        r.second_d = s2 == T_NONE ? (uint16_t)OCB_DIST_INF : (uint16_t)(s2 >> T_SHIFT);
        out[g] = r;
    }
    if (col_out && g < n2)
        col_out[g] = (uint32_t)(~column_key); // zero (nothing seen) -> OCB_NO_INDEX
}


// ---------------------------------------------------------------------------------------------------------------------
// K1T third form: the query operand lives in TENSOR MEMORY (tcgen05.mma with A from TMEM). Measured on the second form
// (ncu, profiles/): the MMA warp stalls on the issue of UTCIMMA itself, an M 128 x N 128 x K 32 MMA takes ~130 cycles and
// an N 64 one ~87 against 64 / 32 of arithmetic -- two cycles per 128-byte shared-memory wavefront of operand fetch
// (A: 32 wavefronts per MMA whatever N is, B: N / 4). With A in TMEM an MMA fetches only its B rows, the whole 227 KB of
// shared memory is the candidate ring (six stages of 64 rows), and nothing has to expand the query rows in global
// memory: the epilogue warps build the s8 image of their own query rows in registers and tcgen05.st it.
// TMEM map (512 columns): [0, 256) A = two query tiles x 128 columns (column c = K bytes 4c .. 4c + 3 of the lane's row),
// [256, 512) two accumulator buffers x (two query tiles x 64 candidate columns).
constexpr int W_STAGES = 6;
constexpr uint32_t W_NT = 64, W_MT = 2, W_QROWS = 256;
constexpr uint32_t W_B_BYTES = W_NT * T_KBYTES;
constexpr uint32_t W_ACC0 = 256;

struct K1T4Params
{
    const uint4 *q_rows;       // the caller's query rows, [n1][4] uint4
    const uint4 *c_tiles;      // expanded candidate tiles of 64 rows
    uint32_t *part;            // [slot * 2 + block][n1_padded][2] packed (s1, s2)
    unsigned long long *col64; // [n2] complemented (distance, query) keys, zeroed by the expansion kernel
    uint32_t n1, n1_padded, n2;
    uint32_t c_steps, total_steps, per;
};

template <bool COL> __global__ void __launch_bounds__(V_THREADS, 1) k1t4_top2_kernel(const K1T4Params P)
{
    constexpr uint32_t G = W_MT * 4;
    extern __shared__ __align__(128) unsigned char k1t4_smem[];
    unsigned char *sB = k1t4_smem;
    uint64_t *bars = reinterpret_cast<uint64_t *>(k1t4_smem + W_STAGES * W_B_BYTES);
    uint64_t *a_full = bars, *d_full = bars + 2, *d_empty = bars + 4, *c_full = bars + 6, *c_empty = bars + 8,
             *b_full = bars + 10, *b_empty = bars + 10 + W_STAGES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 10 + 2 * W_STAGES);
    static_assert((10 + 2 * W_STAGES + 1) * 8 <= V_BAR_BYTES, "barrier block");
    // cross-check keys of a step, [2 buffers][8 row groups][64 candidates] (32-bit words: four keys per 16-byte store)
    uint32_t *colmin = reinterpret_cast<uint32_t *>(k1t4_smem + W_STAGES * W_B_BYTES + V_BAR_BYTES);

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t l_begin = blockIdx.x * P.per;
    const uint32_t l_end = min(P.total_steps, l_begin + P.per);

    if (threadIdx.x == 0)
    {
        mbar_init(a_full, V_EPI_WARPS);
        for (int s = 0; s < 2; s++)
        {
            mbar_init(&d_full[s], 1);
            mbar_init(&d_empty[s], V_EPI_WARPS);
            mbar_init(&c_full[s], V_EPI_WARPS);
            mbar_init(&c_empty[s], 1);
        }
        for (int s = 0; s < W_STAGES; s++)
        {
            mbar_init(&b_full[s], 1);
            mbar_init(&b_empty[s], 1);
        }
        mbar_fence_init();
    }
    if (warp == V_EPI_WARPS + 1)
    {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == V_EPI_WARPS)
    {
        // ---- producer: candidate steps through the ring (the whole warp runs the loop, one elected lane issues)
        uint32_t t = 0;
        for (uint32_t l = l_begin; l < l_end;)
        {
            const uint32_t qb = l / P.c_steps, ct = l - qb * P.c_steps;
            const uint32_t n = min(P.c_steps - ct, l_end - l);
            for (uint32_t i = 0; i < n; i++, t++)
            {
                const uint32_t s = t % W_STAGES;
                if (t >= (uint32_t)W_STAGES)
                    wait_or_trap(&b_empty[s], ((t / W_STAGES) - 1) & 1);
                if (elect_one())
                {
                    mbar_expect_tx(&b_full[s], W_B_BYTES);
                    bulk_g2s(sB + s * W_B_BYTES, P.c_tiles + (size_t)(ct + i) * (W_B_BYTES / 16), W_B_BYTES, &b_full[s]);
                }
                __syncwarp();
            }
            l += n;
        }
    }
    else if (warp == V_EPI_WARPS + 1)
    {
        // ---- MMA issuer: whole warp in the loop, one elected lane issues (see the second form)
        const uint32_t idesc =
            (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(W_NT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t b_addr0 = smem_u32(sB);
        uint32_t t = 0, seg = 0;
        for (uint32_t l = l_begin; l < l_end; seg++)
        {
            const uint32_t qb = l / P.c_steps, ct = l - qb * P.c_steps;
            const uint32_t n = min(P.c_steps - ct, l_end - l);
            wait_or_trap(a_full, seg & 1); // the sixteen epilogue warps have stored this segment's query rows
            for (uint32_t i = 0; i < n; i++, t++)
            {
                const uint32_t s = t % W_STAGES, buf = t & 1;
                wait_or_trap(&b_full[s], (t / W_STAGES) & 1);
                if (t >= 2)
                    wait_or_trap(&d_empty[buf], ((t >> 1) - 1) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (elect_one())
                {
                    const uint32_t b_addr = b_addr0 + s * W_B_BYTES;
#pragma unroll
                    for (uint32_t j = 0; j < T_KBYTES / 32; j++)
                    {
                        const uint64_t bdesc = umma_desc_lbo<(int)W_NT>(b_addr + j * 2 * (W_NT * 16u));
#pragma unroll
                        for (uint32_t m = 0; m < W_MT; m++)
                        {
                            const uint32_t d_tmem = tmem_base + W_ACC0 + buf * 128u + m * W_NT;
                            const uint32_t a_tmem = tmem_base + m * 128u + j * 8u; // K = 32 bytes = 8 columns per MMA
                            const uint32_t accumulate = j > 0 ? 1u : 0u;
                            asm volatile("{\n\t"
                                         ".reg .pred p;\n\t"
                                         "setp.ne.b32 p, %4, 0;\n\t"
                                         "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t"
                                         "}\n" ::"r"(d_tmem),
                                         "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
                                         : "memory");
                        }
                    }
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                                     smem_u32(&b_empty[s]))
                                 : "memory");
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                                     smem_u32(&d_full[buf]))
                                 : "memory");
                }
                __syncwarp();
            }
            l += n;
        }
    }
    else if (warp == V_EPI_WARPS + 2)
    {
        // ---- cross-check flush (as in the second form)
        if constexpr (COL)
        {
            uint32_t t = 0;
            for (uint32_t l = l_begin; l < l_end;)
            {
                const uint32_t qb = l / P.c_steps, ct = l - qb * P.c_steps;
                const uint32_t n = min(P.c_steps - ct, l_end - l);
                for (uint32_t i = 0; i < n; i++, t++)
                {
                    const uint32_t b = t & 1;
                    wait_or_trap(&c_full[b], (t >> 1) & 1);
                    const uint32_t *cm = colmin + b * (G * W_NT);
                    uint32_t best[W_NT / 32];
#pragma unroll
                    for (uint32_t k = 0; k < W_NT / 32; k++)
                    {
                        const uint32_t c = k * 32u + lane;
                        uint32_t m = 0;
#pragma unroll
                        for (uint32_t g = 0; g < G; g++)
                        {
                            const uint32_t key = cm[g * W_NT + c];
                            const uint32_t cand = ((key >> 5) << 8) | (255u - (g * 32u + 31u - (key & 31u)));
                            m = max(m, key ? cand : 0u);
                        }
                        best[k] = m;
                    }
                    __syncwarp();
                    if (lane == 0)
                        mbar_arrive(&c_empty[b]);
#pragma unroll
                    for (uint32_t k = 0; k < W_NT / 32; k++)
                    {
                        const uint32_t pos = (ct + i) * W_NT + k * 32u + lane;
                        if (best[k] != 0u && pos < P.n2)
                        {
                            const unsigned long long key =
                                ((unsigned long long)(513u - (best[k] >> 8)) << 32) |
                                (unsigned long long)(qb * W_QROWS + 255u - (best[k] & 255u));
                            red_max_u64_global(&P.col64[pos], ~key);
                        }
                    }
                }
                l += n;
            }
        }
    }
    else
    {
        // ---- epilogue warps: quarter = warp % 4 (TMEM lanes), block = warp / 4 = (query tile, column half)
        const uint32_t quarter = warp & 3u, blk = warp >> 2;
        const uint32_t mt = blk >> 1, cblk = blk & 1u;
        const uint32_t row_local = mt * 128u + quarter * 32u + lane;
        const uint32_t g = mt * 4u + quarter;
        const uint32_t lane_field = (quarter * 32u) << 16;
        // the half of this thread's query row that this warp stores: K bytes cblk * 256 .. + 256 = bits cblk * 256 ..
        auto load_row = [&](uint32_t qb, uint4 &lo, uint4 &hi) {
            const uint32_t qpos = qb * W_QROWS + row_local;
            lo = hi = make_uint4(0u, 0u, 0u, 0u);
            if (qpos < P.n1)
            {
                lo = P.q_rows[(size_t)qpos * 4 + cblk * 2];
                hi = P.q_rows[(size_t)qpos * 4 + cblk * 2 + 1];
            }
        };
        auto store_a = [&](const uint4 &bits, uint32_t col0) {
            // 128 bits -> 32 columns of four s8 each (+1 for a set bit, -1 for a clear one)
            const uint32_t w[4] = {bits.x, bits.y, bits.z, bits.w};
            uint32_t v[32];
#pragma unroll
            for (int j = 0; j < 32; j++)
            {
                const uint32_t nib = (w[j >> 3] >> ((j & 7) * 4)) & 15u;
                const uint32_t b = (nib * 0x00204081u) & 0x01010101u;
                v[j] = 0xFFFFFFFFu - b * 0xFEu;
            }
            const uint32_t taddr = tmem_base + lane_field + col0;
            asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
                         "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
                         "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n" ::"r"(taddr),
                         "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
                         "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]),
                         "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]),
                         "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
                         : "memory");
        };
        uint4 row_lo, row_hi;
        if (l_begin < l_end)
            load_row(l_begin / P.c_steps, row_lo, row_hi);
        uint32_t t = 0;
        for (uint32_t l = l_begin; l < l_end;)
        {
            const uint32_t qb = l / P.c_steps, ct = l - qb * P.c_steps;
            const uint32_t n = min(P.c_steps - ct, l_end - l);
            const uint32_t qpos = qb * W_QROWS + row_local;
            // This segment's query rows into TMEM. Every MMA of the previous segment has completed: this warp has seen
            // d_full of its last step, which the MMA warp commits after all of them.
            store_a(row_lo, mt * 128u + cblk * 64u);
            store_a(row_hi, mt * 128u + cblk * 64u + 32u);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0)
                mbar_arrive(a_full);
            if (l + n < l_end) // the next segment's rows are on their way while this one is worked on
                load_row(qb + 1, row_lo, row_hi);
            const uint32_t key_mul = qpos < P.n1 ? 16u : 0u;
            const uint32_t key_add = qpos < P.n1 ? (8224u + 31u - lane) : 0u;
            uint32_t s1 = T_NONE, s2 = T_NONE;
            auto step = [&](uint32_t cstep, auto partial) {
                const uint32_t buf = t & 1;
                const uint32_t pos0 = cstep * W_NT + cblk * 32u; // position of this block's first column
                wait_or_trap(&d_full[buf], (t >> 1) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                uint32_t d[32];
                const uint32_t taddr = tmem_base + lane_field + W_ACC0 + buf * 128u + blk * 32u;
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
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0)
                    mbar_arrive(&d_empty[buf]);
                uint32_t *cm = colmin + (t & 1u) * (G * W_NT) + g * W_NT + cblk * 32u;
                if constexpr (COL)
                    if (t >= 2)
                        wait_or_trap(&c_empty[t & 1], ((t >> 1) - 1) & 1);
                // Two columns per register: key16 = distance << 5 | column within the block = 8192 + i - 16 * dot (15 bits),
                // even columns in the low halves, odd ones in the high halves; the two smallest per half with the packed
                // 16-bit minimum / maximum instructions (VIMNMX.U16x2, VIMNMX3.U16x2): 1.25 operations per value.
                uint32_t b1 = 0xFFFFFFFFu, b2 = 0xFFFFFFFFu;
                auto packed = [&](int j) {
                    // (dot_odd << 16) + dot_even, then one multiply-add for both keys (the low key is never negative, so
                    // nothing is borrowed from the high half)
                    const uint32_t both = d[2 * j + 1] * 65536u + d[2 * j];
                    uint32_t k = both * (0u - 16u) + (((8192u + 2u * j + 1u) << 16) | (8192u + 2u * j));
                    if constexpr (decltype(partial)::value) // only the last candidate step can hold padding columns
                    {
                        k |= pos0 + 2u * j < P.n2 ? 0u : 0x0000FFFFu;
                        k |= pos0 + 2u * j + 1u < P.n2 ? 0u : 0xFFFF0000u;
                    }
                    return k;
                };
#pragma unroll
                for (int i = 0; i < 32; i += 4)
                {
                    const uint32_t p = packed(i / 2), q = packed(i / 2 + 1);
                    const uint32_t lo = __vminu2(p, q), hi = __vmaxu2(p, q);
                    b2 = __vimin3_u16x2(__vmaxu2(b1, lo), b2, hi);
                    b1 = __vminu2(b1, lo);
                    if constexpr (COL)
                    {
                        // the warp's best (distance, row) for the candidates pos0 + i .. i + 3: four keys, one 16-byte
                        // store (every lane stores the same words)
                        uint32_t m[4];
#pragma unroll
                        for (int k = 0; k < 4; k++)
                            m[k] = __reduce_max_sync(0xFFFFFFFFu, d[i + k] * key_mul + key_add);
                        *reinterpret_cast<uint4 *>(cm + i) = make_uint4(m[0], m[1], m[2], m[3]);
                    }
                }
                if constexpr (COL)
                {
                    __syncwarp();
                    if (lane == 0)
                        mbar_arrive(&c_full[t & 1]);
                }
                // the block's two smallest: even and odd columns together, then rebased to v = distance << 20 | position
                const uint32_t e1 = b1 & 0xFFFFu, o1 = b1 >> 16, e2 = b2 & 0xFFFFu, o2 = b2 >> 16;
                const uint32_t m1 = min(e1, o1);
                const uint32_t m2 = (uint32_t)__vimin3_s32((int)max(e1, o1), (int)e2, (int)o2);
                auto rebase = [&](uint32_t k) {
                    // (k >> 5) << 20 | (pos0 + (k & 31))
                    const uint32_t v = k * 32768u + pos0 - (k & 31u) * 32767u;
                    if constexpr (decltype(partial)::value)
                        return k == 0xFFFFu ? T_NONE : v;
                    else
                        return v;
                };
                const uint32_t g1 = rebase(m1), g2 = rebase(m2);
                s2 = (uint32_t)min(min(max(s1, g1), s2), g2);
                s1 = min(s1, g1);
            };
            for (uint32_t i = 0; i < n; i++, t++)
            {
                if ((ct + i + 1) * W_NT <= P.n2)
                    step(ct + i, std::false_type());
                else
                    step(ct + i, std::true_type());
            }
            const uint32_t slot = blockIdx.x - (qb * P.c_steps) / P.per;
            uint32_t *out = P.part + ((size_t)(slot * 2u + cblk) * P.n1_padded + qpos) * 2;
            out[0] = s1, out[1] = s2;
            l += n;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == V_EPI_WARPS + 1)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
}

struct K1T2Layout
{
    uint32_t q_groups, c_steps, n1p, total, per, ctas, max_slots;
    size_t o_qt, o_ct, o_part, o_col, bytes;
};

K1T2Layout k1t2_layout(size_t n1, size_t n2, bool col, int sms, int mt, int nt)
{
    K1T2Layout L;
    const uint32_t qrows = (uint32_t)mt * 128u;
    L.q_groups = (uint32_t)((n1 + qrows - 1) / qrows);
    L.c_steps = (uint32_t)((n2 + nt - 1) / nt);
    L.n1p = L.q_groups * qrows;
    L.total = L.q_groups * L.c_steps;
    const uint32_t want = std::max(1u, std::min<uint32_t>((uint32_t)std::max(sms, 1), L.total));
    L.per = (L.total + want - 1) / want;
    L.ctas = (L.total + L.per - 1) / L.per;
    // spans that touch one query tile group: at most ceil(c_steps / per) + 1
    L.max_slots = (L.c_steps + L.per - 1) / L.per + 1;
    size_t off = 0;
    auto take = [&off](size_t bytes) {
        const size_t o = off;
        off = (off + bytes + 255) / 256 * 256;
        return o;
    };
    L.o_qt = take((size_t)L.n1p * T_KBYTES);
    L.o_ct = take((size_t)L.c_steps * nt * T_KBYTES);
    L.o_part = take((size_t)L.max_slots * (nt / 32) * L.n1p * 2 * sizeof(uint32_t));
    L.o_col = take(col ? n2 * sizeof(unsigned long long) : 0);
    L.bytes = off;
    return L;
}

template <bool COL, int MT, int NT, int STAGES>
int k1t2_run(const K1T2Params &P, const K1T2Layout &L, cudaStream_t stream)
{
    constexpr size_t smem = (size_t)MT * V_A_TILE_BYTES + (size_t)STAGES * NT * T_KBYTES + V_BAR_BYTES + V_COLMIN_BYTES;
    static_assert(smem <= 232448, "227 KB of dynamic shared memory per CTA");
    OCB_CUDA(cudaFuncSetAttribute(k1t2_top2_kernel<COL, MT, NT, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
    k1t2_top2_kernel<COL, MT, NT, STAGES><<<L.ctas, V_THREADS, smem, stream>>>(P);
    return 0;
}

template <bool COL> int k1t4_run(const K1T4Params &P, uint32_t ctas, cudaStream_t stream)
{
    constexpr size_t smem = (size_t)W_STAGES * W_B_BYTES + V_BAR_BYTES + 2 * V_COLMIN_BYTES;
    static_assert(smem <= 232448, "227 KB of dynamic shared memory per CTA");
    OCB_CUDA(cudaFuncSetAttribute(k1t4_top2_kernel<COL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k1t4_top2_kernel<COL><<<ctas, V_THREADS, smem, stream>>>(P);
    return 0;
}

int k1t2_launch(bool a_in_tmem, int mt, int nt, const void *d_q, size_t n1, const void *d_c, size_t n2, ocb_top2 *d_out,
                uint32_t *d_col_best_q, void *d_workspace, int sms, cudaStream_t stream)
{
    const bool col = d_col_best_q != nullptr;
    const K1T2Layout L = k1t2_layout(n1, n2, col, sms, mt, nt);
    char *ws = static_cast<char *>(d_workspace);
    uint4 *qt = reinterpret_cast<uint4 *>(ws + L.o_qt), *ct = reinterpret_cast<uint4 *>(ws + L.o_ct);
    K1T2Params P;
    P.q_tiles = qt, P.c_tiles = ct;
    P.part = reinterpret_cast<uint32_t *>(ws + L.o_part);
    P.col64 = col ? reinterpret_cast<unsigned long long *>(ws + L.o_col) : nullptr;
    P.n1 = (uint32_t)n1, P.n1_padded = L.n1p, P.n2 = (uint32_t)n2;
    P.c_steps = L.c_steps, P.total_steps = L.total, P.per = L.per;
    // third form: the query rows are expanded inside the search kernel (registers -> tensor memory)
    const uint32_t slabs_a = a_in_tmem ? 0u : L.n1p / 128u, slabs_b = (L.c_steps * (uint32_t)nt + 127u) / 128u;
    k1t2_expand_kernel<<<dim3(slabs_a + slabs_b, 4), 128, 0, stream>>>(
        static_cast<const uint4 *>(d_q), (uint32_t)n1, slabs_a, qt, static_cast<const uint4 *>(d_c), (uint32_t)n2,
        (uint32_t)nt, L.c_steps * (uint32_t)nt, ct, P.col64, (uint32_t)n2);
    int rc;
    if (a_in_tmem)
    {
        K1T4Params Q;
        Q.q_rows = static_cast<const uint4 *>(d_q), Q.c_tiles = ct, Q.part = P.part, Q.col64 = P.col64;
        Q.n1 = P.n1, Q.n1_padded = P.n1_padded, Q.n2 = P.n2;
        Q.c_steps = P.c_steps, Q.total_steps = P.total_steps, Q.per = P.per;
        rc = col ? k1t4_run<true>(Q, L.ctas, stream) : k1t4_run<false>(Q, L.ctas, stream);
    }
    else if (mt == 1)
        rc = col ? k1t2_run<true, 1, 128, 2>(P, L, stream) : k1t2_run<false, 1, 128, 2>(P, L, stream);
    else
        rc = col ? k1t2_run<true, 2, 64, 3>(P, L, stream) : k1t2_run<false, 2, 64, 3>(P, L, stream);
    if (rc)
        return rc;
    const uint32_t finish = (uint32_t)std::max(n1, col ? n2 : (size_t)0);
    k1t2_finish_kernel<<<(finish + 255) / 256, 256, 0, stream>>>(P.part, L.n1p, (uint32_t)n1, (uint32_t)mt * 128u,
                                                               (uint32_t)nt / 32u, L.c_steps, L.per, d_out, P.col64,
                                                               (uint32_t)n2, d_col_best_q);
    count_launch(3);
    OCB_CUDA(cudaGetLastError());
    return 0;
}

// ---- diagnostic (include/ocb_probe.h): how fast the tensor pipe retires tcgen05.mma.kind::i8 (M 128 x N x K 32) as a
// function of N, of the number of independent accumulators the MMAs rotate over, and of where A lives
__global__ void __launch_bounds__(128, 1) umma_probe_kernel(int a_in_tmem, int n, int chains, int rounds,
                                                            unsigned long long *cycles_out)
{
    extern __shared__ __align__(128) unsigned char probe_smem[];
    unsigned char *sA = probe_smem;                 // 64 KB: one query tile (shared-memory A only)
    unsigned char *sB = probe_smem + 65536;         // up to 256 rows x 512 B
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (uint32_t i = threadIdx.x; i < (65536u + 131072u) / 16; i += blockDim.x)
        reinterpret_cast<uint4 *>(probe_smem)[i] = make_uint4(0x01FF01FFu, 0xFF01FF01u, 0x0101FFFFu, 0xFFFF0101u);
    if (threadIdx.x == 0)
    {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    fence_proxy_async_smem();
    if (threadIdx.x < 32)
    {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = slot;
    if (threadIdx.x < 32)
    {
        const uint32_t idesc =
            (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t a_addr = smem_u32(sA), b_addr = smem_u32(sB);
        const uint32_t lbo_b = (uint32_t)n * 16u;
        long long t0 = 0, t1 = 0;
        if (elect_one())
        {
            t0 = clock64();
            for (int r = 0; r < rounds; r++)
            {
                const uint32_t j = (uint32_t)r & 15u;
                uint64_t bdesc = (uint64_t)(((b_addr + j * 2 * lbo_b) >> 4) & 0x3FFFu);
                bdesc |= (uint64_t)((lbo_b >> 4) & 0x3FFFu) << 16;
                bdesc |= (uint64_t)((T_SBO >> 4) & 0x3FFFu) << 32;
                bdesc |= 1ull << 46;
                const uint64_t adesc = umma_desc_lbo<128>(a_addr + j * 2 * (128u * 16u));
                const uint32_t a_tmem = tmem_base + 384u + j * 8u;
                for (int c = 0; c < chains; c++)
                {
                    const uint32_t d_tmem = tmem_base + (uint32_t)(c * n);
                    if (a_in_tmem)
                        asm volatile("{\n\t"
                                     ".reg .pred p;\n\t"
                                     "setp.ne.b32 p, %4, 0;\n\t"
                                     "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t"
                                     "}\n" ::"r"(d_tmem),
                                     "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(1u), "r"(0u)
                                     : "memory");
                    else
                        asm volatile("{\n\t"
                                     ".reg .pred p;\n\t"
                                     "setp.ne.b32 p, %4, 0;\n\t"
                                     "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
                                     "}\n" ::"r"(d_tmem),
                                     "l"(adesc), "l"(bdesc), "r"(idesc), "r"(1u), "r"(0u)
                                     : "memory");
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar))
                         : "memory");
            wait_or_trap(&bar, 0);
            t1 = clock64();
            cycles_out[blockIdx.x] = (unsigned long long)(t1 - t0);
        }
        __syncwarp();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
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
    // (and over the forms of the search kernel, so that an option change between the size query and the call is safe)
    size_t worst = 0;
    for (int sms : {148, 132, 108, 64, 1})
    {
        worst = std::max(worst, k1t_layout(n1, n2, col, sms).total);
        worst = std::max(worst, k1t2_layout(n1, n2, col, sms, 1, 128).bytes);
        worst = std::max(worst, k1t2_layout(n1, n2, col, sms, 2, 64).bytes);
    }
    return worst + 256;
}

int k1t_launch(const void *d_q, size_t n1, const void *d_c, size_t n2, ocb_top2 *d_out, uint32_t *d_col_best_q,
               void *d_workspace, int sms, cudaStream_t stream)
{
    // form of the search kernel: 1 = first form (eight epilogue warps, one CTA per (query tile, candidate range)),
    // 2 / 3 = second form with one / two query tiles per CTA, 4 = third form (query operand in tensor memory);
    // 0 = the default (the third form)
    int variant = options().k1t_variant;
    if (variant <= 0 || variant > 4)
        variant = K1T_DEFAULT_VARIANT;
    if (variant == 2)
        return k1t2_launch(false, 1, 128, d_q, n1, d_c, n2, d_out, d_col_best_q, d_workspace, sms, stream);
    if (variant == 3)
        return k1t2_launch(false, 2, 64, d_q, n1, d_c, n2, d_out, d_col_best_q, d_workspace, sms, stream);
    if (variant == 4)
        return k1t2_launch(true, 2, 64, d_q, n1, d_c, n2, d_out, d_col_best_q, d_workspace, sms, stream);
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

int umma_probe(int a_in_tmem, int n, int chains, int rounds, int ctas, double *cycles_per_mma)
{
    if (n < 16 || n > 256 || n % 16 || chains < 1 || chains * n > (a_in_tmem ? 384 : 512) || rounds < 1 || ctas < 1 ||
        !cycles_per_mma)
        return fail_invalid("umma probe arguments");
    unsigned long long *d = nullptr;
    OCB_CUDA(cudaMalloc(&d, sizeof(unsigned long long) * ctas));
    const size_t smem = 65536 + 131072;
    cudaError_t e = cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess)
    {
        umma_probe_kernel<<<ctas, 128, smem>>>(a_in_tmem, n, chains, rounds, d);
        e = cudaDeviceSynchronize();
    }
    unsigned long long worst = 0;
    if (e == cudaSuccess)
    {
        std::vector<unsigned long long> h(ctas);
        e = cudaMemcpy(h.data(), d, sizeof(unsigned long long) * ctas, cudaMemcpyDeviceToHost);
        for (unsigned long long v : h)
            worst = std::max(worst, v);
    }
    cudaFree(d);
    OCB_CUDA(e);
    *cycles_per_mma = (double)worst / ((double)rounds * chains);
    return 0;
}

} // namespace ocb

extern "C" int ocb_probe_umma(int a_in_tmem, int n, int chains, int rounds, int ctas, double *cycles_per_mma)
{
    return ocb::umma_probe(a_in_tmem, n, chains, rounds, ctas, cycles_per_mma);
}
