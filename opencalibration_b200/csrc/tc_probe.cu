// Tensor-core probe for the Hamming top-2 search (include/ocb_probe.h: ocb_probe_tensor_top2). NOT the product path:
// K1 (hamming_top2.cu) computes distances as XOR + POPC on the integer pipes, as north_star specifies. This file is the
// one measured experiment behind that choice.
//
// Formulation. With bits mapped to s = 2 b - 1 in {-1, +1}, the dot product of two rows over 512 positions is
// 512 - 2 * hamming (equal bits contribute +1, different bits -1; the 26 padding bits are equal), i.e. an EXACT dense
// integer contraction: tcgen05.mma.kind::i8 (s8 x s8 -> s32 in TMEM), M = 128 queries x N = 128 candidates x K = 512.
//   * tc_expand_kernel: 64-byte descriptor rows -> s8 tiles of 128 rows in the K-major no-swizzle ("interleave")
//     canonical layout the UMMA shared-memory descriptor addresses: [k chunk of 16 B][row][16 B] (core matrix = 8 rows x
//     16 B contiguous, SBO = 128 B between 8-row groups, LBO = 2048 B between K chunks), so ONE 64 KB bulk copy per
//     tile fills shared memory;
//   * tc_top2_kernel: one CTA per (query tile, candidate range), 6 warps: warp 0 = bulk-copy producer (query tile once,
//     candidate tiles through a 2-stage ring), warp 1 = the single thread that issues 16 MMAs (K = 32 each) per
//     candidate tile into one of two 128-column TMEM accumulators and commits them to mbarriers, warps 2-5 = epilogue:
//     tcgen05.ld of 32 columns at a time (thread = query row), v = distance << 20 | position with one IMAD from the
//     dot product, the same branch-free two-smallest update as K1;
//   * tc_merge_kernel: combines the candidate ranges of a query (global positions, so the packed values merge).
// Every wait is bounded: a barrier that does not flip within 4 M polls traps instead of hanging the device.
#include "ocb_internal.cuh"

#include "../../include/ocb_probe.h"

#include <algorithm>
#include <vector>

namespace ocb
{
namespace
{
constexpr int TC_ROWS = 128;                        // rows per tile (M and N of one MMA)
constexpr int TC_KBYTES = 512;                      // s8 elements per row
constexpr uint32_t TC_TILE_BYTES = TC_ROWS * TC_KBYTES; // 64 KB
constexpr uint32_t TC_LBO = TC_ROWS * 16;           // bytes between K chunks
constexpr uint32_t TC_SBO = 8 * 16;                 // bytes between 8-row groups
constexpr int TC_STAGES = 2;
constexpr int TC_THREADS = 192;
constexpr uint32_t TC_SHIFT = 20;
constexpr uint32_t TC_NONE = 0xFFFFFFFFu;
constexpr uint32_t TC_TMEM_COLS = 256;

// rows [n][8] u64 -> tiles [ceil(n/128)][32 k chunks][128 rows][16 B] of s8 (+1 / -1); rows past n: all -1
__global__ void __launch_bounds__(256) tc_expand_kernel(const uint64_t *__restrict__ rows, uint32_t n, uint32_t n_padded,
                                                         uint4 *__restrict__ tiles)
{
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; // (row, k chunk)
    if (g >= n_padded * 32u)
        return;
    const uint32_t kc = g & 31u, row = g >> 5;
    uint32_t bits16 = 0;
    if (row < n)
        bits16 = (uint32_t)(rows[(size_t)row * 8 + (kc >> 2)] >> ((kc & 3u) * 16u)) & 0xFFFFu;
    auto spread = [](uint32_t nib) { // 4 bits -> 4 bytes of 0x01 (bit set) / 0xFF (bit clear)
        const uint32_t b = (nib * 0x00204081u) & 0x01010101u;
        return 0xFFFFFFFFu - b * 0xFEu;
    };
    uint4 v;
    v.x = spread(bits16 & 15u), v.y = spread((bits16 >> 4) & 15u), v.z = spread((bits16 >> 8) & 15u),
    v.w = spread((bits16 >> 12) & 15u);
    const uint32_t tile = row / TC_ROWS, r = row % TC_ROWS;
    tiles[((size_t)tile * 32 + kc) * TC_ROWS + r] = v;
}

__device__ __forceinline__ void bounded_wait(uint64_t *bar, uint32_t parity, uint32_t *err, uint32_t code)
{
    for (uint32_t spin = 0; spin < (1u << 22); spin++)
        if (mbar_try_wait(bar, parity))
            return;
    atomicExch(err, code);
    __threadfence_system();
    asm volatile("trap;");
}

__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr)
{
    // cute::UMMA::SmemDescriptor: start address, leading / stride byte offsets (all >> 4), version 1, no swizzle
    uint64_t d = (uint64_t)((smem_addr >> 4) & 0x3FFFu);
    d |= (uint64_t)((TC_LBO >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((TC_SBO >> 4) & 0x3FFFu) << 32;
    d |= 1ull << 46;
    return d;
}

struct TcParams
{
    const uint4 *q_tiles; // expanded query tiles
    const uint4 *c_tiles; // expanded candidate tiles
    uint32_t *part;       // [splits][n1_padded][2] packed (s1, s2)
    uint32_t *err;
    uint32_t n1_padded, n2, c_tiles_total, tiles_per_split;
};

__global__ void __launch_bounds__(TC_THREADS, 1) tc_top2_kernel(const TcParams P)
{
    extern __shared__ __align__(128) unsigned char tc_smem[];
    unsigned char *sA = tc_smem;
    unsigned char *sB = tc_smem + TC_TILE_BYTES;
    uint64_t *bars = reinterpret_cast<uint64_t *>(tc_smem + (1 + TC_STAGES) * TC_TILE_BYTES);
    uint64_t *a_full = bars, *b_full = bars + 1, *b_empty = bars + 3, *d_full = bars + 5, *d_empty = bars + 7;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 9);

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t qtile = blockIdx.x, split = blockIdx.y;
    const uint32_t t_begin = split * P.tiles_per_split;
    const uint32_t t_end = min(P.c_tiles_total, t_begin + P.tiles_per_split);
    const uint32_t ntiles = t_end > t_begin ? t_end - t_begin : 0;

    if (threadIdx.x == 0)
    {
        mbar_init(a_full, 1);
        for (int s = 0; s < TC_STAGES; s++)
        {
            mbar_init(&b_full[s], 1);
            mbar_init(&b_empty[s], 1);
            mbar_init(&d_full[s], 1);
            mbar_init(&d_empty[s], 4);
        }
        mbar_fence_init();
    }
    if (warp == 1)
    {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(TC_TMEM_COLS));
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
            mbar_expect_tx(a_full, TC_TILE_BYTES);
            bulk_g2s(sA, P.q_tiles + (size_t)qtile * (TC_TILE_BYTES / 16), TC_TILE_BYTES, a_full);
            for (uint32_t t = 0; t < ntiles; t++)
            {
                const uint32_t s = t % TC_STAGES;
                if (t >= TC_STAGES)
                    bounded_wait(&b_empty[s], ((t / TC_STAGES) - 1) & 1, P.err, 1);
                mbar_expect_tx(&b_full[s], TC_TILE_BYTES);
                bulk_g2s(sB + s * TC_TILE_BYTES, P.c_tiles + (size_t)(t_begin + t) * (TC_TILE_BYTES / 16), TC_TILE_BYTES,
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
            const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_ROWS >> 3) << 17) |
                                   ((uint32_t)(TC_ROWS >> 4) << 24);
            bounded_wait(a_full, 0, P.err, 2);
            const uint32_t a_addr = smem_u32(sA);
            for (uint32_t t = 0; t < ntiles; t++)
            {
                const uint32_t s = t % TC_STAGES, buf = t & 1;
                bounded_wait(&b_full[s], (t / TC_STAGES) & 1, P.err, 3);
                if (t >= 2)
                    bounded_wait(&d_empty[buf], ((t >> 1) - 1) & 1, P.err, 4);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t b_addr = smem_u32(sB + s * TC_TILE_BYTES);
                const uint32_t d_tmem = tmem_base + buf * TC_ROWS;
#pragma unroll
                for (uint32_t j = 0; j < TC_KBYTES / 32; j++)
                {
                    const uint64_t adesc = umma_desc(a_addr + j * 2 * TC_LBO), bdesc = umma_desc(b_addr + j * 2 * TC_LBO);
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
        const uint32_t quarter = warp & 3u;             // the TMEM lanes a warp may read: 32 * (warp id % 4) ..
        const uint32_t row = quarter * 32u + lane;       // query row of this thread within the tile
        uint32_t s1 = TC_NONE, s2 = TC_NONE;
        for (uint32_t t = 0; t < ntiles; t++)
        {
            const uint32_t buf = t & 1;
            bounded_wait(&d_full[buf], (t >> 1) & 1, P.err, 5);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t pos0 = (t_begin + t) * TC_ROWS;
#pragma unroll 1
            for (uint32_t c0 = 0; c0 < TC_ROWS; c0 += 32)
            {
                uint32_t d[32];
                const uint32_t taddr = tmem_base + ((quarter * 32u) << 16) + buf * TC_ROWS + c0;
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
                const uint32_t base = (512u << (TC_SHIFT - 1)) + pos0 + c0; // v = (512 - dot) / 2 << 20 | position
#pragma unroll
                for (int i = 0; i < 32; i++)
                {
                    uint32_t v = base + (uint32_t)i - (d[i] << (TC_SHIFT - 1)); // dot is even: (512 - dot) << 19 is exact
                    v = pos0 + c0 + (uint32_t)i < P.n2 ? v : TC_NONE;
                    s2 = min(s2, max(s1, v));
                    s1 = min(s1, v);
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0)
                mbar_arrive(&d_empty[buf]);
        }
        uint32_t *out = P.part + ((size_t)split * P.n1_padded + (size_t)qtile * TC_ROWS + row) * 2;
        out[0] = s1, out[1] = s2;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TC_TMEM_COLS));
}

__global__ void __launch_bounds__(256) tc_merge_kernel(const uint32_t *__restrict__ part, uint32_t splits,
                                                        uint32_t n1_padded, uint32_t n1, ocb_top2 *__restrict__ out)
{
    const uint32_t qi = blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= n1)
        return;
    uint32_t s1 = TC_NONE, s2 = TC_NONE;
    for (uint32_t s = 0; s < splits; s++)
    {
        const uint32_t *p = part + ((size_t)s * n1_padded + qi) * 2;
        const uint32_t a = p[0], b = p[1];
        s2 = min(s2, max(s1, a));
        s1 = min(s1, a);
        s2 = min(s2, max(s1, b));
        s1 = min(s1, b);
    }
    ocb_top2 r;
    r.best_k = s1 == TC_NONE ? 0u : (s1 & ((1u << TC_SHIFT) - 1u));
    r.best_d = s1 == TC_NONE ? (uint16_t)OCB_DIST_INF : (uint16_t)(s1 >> TC_SHIFT);
    r.second_d = s2 == TC_NONE ? (uint16_t)OCB_DIST_INF : (uint16_t)(s2 >> TC_SHIFT);
    out[qi] = r;
}
} // namespace
} // namespace ocb

using namespace ocb;

extern "C" int ocb_probe_tensor_top2(const uint64_t *q, size_t n1, const uint64_t *c, size_t n2, ocb_top2 *out, int reps,
                                     double *ms3)
{
    if (!q || !c || !out || n1 == 0 || n2 == 0 || n2 >= (1u << TC_SHIFT) || n1 >= (1u << 24))
        return fail_invalid("ocb_probe_tensor_top2: sizes");
    int dev = 0;
    OCB_CUDA(cudaGetDevice(&dev));
    const uint32_t q_tiles = (uint32_t)((n1 + TC_ROWS - 1) / TC_ROWS), c_tiles = (uint32_t)((n2 + TC_ROWS - 1) / TC_ROWS);
    const uint32_t n1p = q_tiles * TC_ROWS, n2p = c_tiles * TC_ROWS;
    // candidate ranges: enough CTAs for whole waves of one CTA per SM
    const int sms = sm_count(dev);
    uint32_t splits = 1;
    double best = 1e30;
    for (uint32_t s = 1; s <= std::min<uint32_t>(c_tiles, 32); s++)
    {
        const uint32_t per = (c_tiles + s - 1) / s, ctas = q_tiles * ((c_tiles + per - 1) / per);
        const double waves = (double)((ctas + sms - 1) / sms);
        const double cost = waves * (per + 1.5); // tiles per CTA + the query tile load and pipeline fill
        if (cost < best)
            best = cost, splits = (c_tiles + per - 1) / per;
    }
    const uint32_t per = (c_tiles + splits - 1) / splits;
    splits = (c_tiles + per - 1) / per;
    void *d_q = nullptr, *d_c = nullptr, *d_qt = nullptr, *d_ct = nullptr, *d_part = nullptr, *d_out = nullptr, *d_err = nullptr;
    cudaStream_t st = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    int rc = 0;
    auto cleanup = [&]() {
        for (void *p : {d_q, d_c, d_qt, d_ct, d_part, d_out, d_err})
            if (p)
                cudaFree(p);
        for (cudaEvent_t e : ev)
            if (e)
                cudaEventDestroy(e);
        if (st)
            cudaStreamDestroy(st);
    };
#define TC_TRY(call)                                                                                                   \
    do                                                                                                                 \
    {                                                                                                                  \
        cudaError_t e__ = (call);                                                                                      \
        if (e__ != cudaSuccess)                                                                                        \
        {                                                                                                              \
            rc = fail_cuda(e__, #call, __FILE__, __LINE__);                                                            \
            cleanup();                                                                                                 \
            return rc;                                                                                                 \
        }                                                                                                              \
    } while (0)
    TC_TRY(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    for (cudaEvent_t &e : ev)
        TC_TRY(cudaEventCreate(&e));
    TC_TRY(cudaMalloc(&d_q, n1 * 64));
    TC_TRY(cudaMalloc(&d_c, n2 * 64));
    TC_TRY(cudaMalloc(&d_qt, (size_t)q_tiles * TC_TILE_BYTES));
    TC_TRY(cudaMalloc(&d_ct, (size_t)c_tiles * TC_TILE_BYTES));
    TC_TRY(cudaMalloc(&d_part, (size_t)splits * n1p * 2 * sizeof(uint32_t)));
    TC_TRY(cudaMalloc(&d_out, n1 * sizeof(ocb_top2)));
    TC_TRY(cudaMalloc(&d_err, sizeof(uint32_t)));
    TC_TRY(cudaMemsetAsync(d_err, 0, sizeof(uint32_t), st));
    TC_TRY(cudaMemcpyAsync(d_q, q, n1 * 64, cudaMemcpyHostToDevice, st));
    TC_TRY(cudaMemcpyAsync(d_c, c, n2 * 64, cudaMemcpyHostToDevice, st));
    const size_t smem = (size_t)(1 + TC_STAGES) * TC_TILE_BYTES + 128;
    TC_TRY(cudaFuncSetAttribute(tc_top2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    TcParams P;
    P.q_tiles = static_cast<const uint4 *>(d_qt), P.c_tiles = static_cast<const uint4 *>(d_ct);
    P.part = static_cast<uint32_t *>(d_part), P.err = static_cast<uint32_t *>(d_err);
    P.n1_padded = n1p, P.n2 = (uint32_t)n2, P.c_tiles_total = c_tiles, P.tiles_per_split = per;
    if (reps < 1)
        reps = 1;
    for (int r = 0; r <= reps; r++) // round 0 = warm-up
    {
        if (r == 1)
            TC_TRY(cudaEventRecord(ev[0], st));
        tc_expand_kernel<<<(n1p * 32 + 255) / 256, 256, 0, st>>>(static_cast<const uint64_t *>(d_q), (uint32_t)n1, n1p,
                                                                static_cast<uint4 *>(d_qt));
        tc_expand_kernel<<<(n2p * 32 + 255) / 256, 256, 0, st>>>(static_cast<const uint64_t *>(d_c), (uint32_t)n2, n2p,
                                                                static_cast<uint4 *>(d_ct));
        count_launch(2);
    }
    TC_TRY(cudaEventRecord(ev[1], st));
    for (int r = 0; r <= reps; r++)
    {
        if (r == 1)
            TC_TRY(cudaEventRecord(ev[2], st));
        tc_top2_kernel<<<dim3(q_tiles, splits), TC_THREADS, smem, st>>>(P);
        tc_merge_kernel<<<((uint32_t)n1 + 255) / 256, 256, 0, st>>>(static_cast<const uint32_t *>(d_part), splits, n1p,
                                                                    (uint32_t)n1, static_cast<ocb_top2 *>(d_out));
        count_launch(2);
    }
    TC_TRY(cudaEventRecord(ev[3], st));
    TC_TRY(cudaGetLastError());
    TC_TRY(cudaMemcpyAsync(out, d_out, n1 * sizeof(ocb_top2), cudaMemcpyDeviceToHost, st));
    TC_TRY(cudaStreamSynchronize(st));
    uint32_t err = 0;
    TC_TRY(cudaMemcpy(&err, d_err, sizeof err, cudaMemcpyDeviceToHost));
    if (ms3)
    {
        float a = 0, b = 0;
        TC_TRY(cudaEventElapsedTime(&a, ev[0], ev[1]));
        TC_TRY(cudaEventElapsedTime(&b, ev[2], ev[3]));
        ms3[0] = a / reps; // expansion of both sets
        ms3[1] = b / reps; // MMA + top-2 + merge
        ms3[2] = (double)splits;
    }
    cleanup();
#undef TC_TRY
    if (err)
    {
        set_last_error("ocb_probe_tensor_top2: a barrier wait timed out (code " + std::to_string(err) + ")");
        return OCB_E_INVALID;
    }
    return 0;
}
