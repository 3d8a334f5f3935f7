// K2/K3 -- hypotheses x correspondences MSAC scoring in IEEE double (sm_100a).
//
// Replaces the score loop of ransac<Model> (reference src/model_inliers/ransac.cpp:183-196, without the SPRT
// early exit) and Model::evaluate (homography_model.cpp:99-118, essential_matrix_model.cpp:91-110,
// fundamental_matrix_model.cpp:89-108) for many hypotheses at once.
//
// Exactness: every multiply / add / divide / sqrt is issued through the *_rn intrinsics, which ptxas never
// contracts into FMAs, in the canonical operation order of the residuals (dot products left to right, see
// h_residual / epi_residual). The MSAC sum of a hypothesis is accumulated by ONE thread sequentially in
// evaluation order, so the result is the reference's left-to-right double sum bit for bit, not a tree sum.
//
// Mapping: a CTA owns K2_HG hypotheses (their matrices sit in shared memory and are read as broadcasts).
// Warps 0..6 ("compute") each take 32 consecutive evaluation positions per round and, for every hypothesis
// of the group, compute the residual, ballot the inlier mask and park the MSAC contribution in a
// double-buffered shared-memory slab (+0.0 for an outlier). Warp 7 ("sum") runs one round behind: lane g adds
// hypothesis g's rows in order, 32 positions at a time (words without an inlier are skipped). The reference adds
// nothing for an outlier; adding +0.0 to a score that is never -0.0 gives the same bits, and it turns the sum into
// straight chains of DADDs whose loads do not depend on the data (a loop over the set bits of the masks made the
// sum warp the critical path). The warps of a CTA synchronise through mbarriers on the two slabs only.
#include "ocb_internal.cuh"
#include "exact_math.cuh"

#include <atomic>
#include <cfloat>

namespace ocb
{

constexpr int K2_HG = 8;          // hypotheses per CTA (upper bound; k2_score picks 4..8 per launch to balance the SMs)
constexpr int K2_CW = 7;          // compute warps
constexpr int K2_TP = K2_CW * 32; // evaluation positions per round
constexpr int K2_THREADS = (K2_CW + 1) * 32;
constexpr int K2_SUBS = 4;     // hypothesis groups per CTA of k2_score_kernel (one 1024-thread CTA per SM)
constexpr int K2_LOCKSTEP = 8; // rounds between the CTA-wide barriers that keep those groups together

// homography_model::error (homography_model.cpp:89-97) on pre-divided coordinates, plain IEEE intrinsics.
// M = [H (9, column-major) | H^-1 (9)]. Out of line: only reached when residual_fast's range test fails.
__device__ __noinline__ double h_residual_ieee(const double *__restrict__ M, double x1, double y1, double x2, double y2)
{
    const double px = __dadd_rn(__dadd_rn(__dmul_rn(M[0], x1), __dmul_rn(M[3], y1)), M[6]);
    const double py = __dadd_rn(__dadd_rn(__dmul_rn(M[1], x1), __dmul_rn(M[4], y1)), M[7]);
    const double pz = __dadd_rn(__dadd_rn(__dmul_rn(M[2], x1), __dmul_rn(M[5], y1)), M[8]);
    const double dx = __dsub_rn(__ddiv_rn(px, pz), x2);
    const double dy = __dsub_rn(__ddiv_rn(py, pz), y2);
    const double fwd = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
    const double qx = __dadd_rn(__dadd_rn(__dmul_rn(M[9], x2), __dmul_rn(M[12], y2)), M[15]);
    const double qy = __dadd_rn(__dadd_rn(__dmul_rn(M[10], x2), __dmul_rn(M[13], y2)), M[16]);
    const double qz = __dadd_rn(__dadd_rn(__dmul_rn(M[11], x2), __dmul_rn(M[14], y2)), M[17]);
    const double ex = __dsub_rn(__ddiv_rn(qx, qz), x1);
    const double ey = __dsub_rn(__ddiv_rn(qy, qz), y1);
    const double bwd = __dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey));
    return __dsqrt_rn(__dmul_rn(__dadd_rn(fwd, bwd), 0.5)); // x/2.0 == x*0.5 exactly (both correctly rounded)
}

// essential_matrix_model::error == fundamental_matrix_model::error
// (essential_matrix_model.cpp:112-123, fundamental_matrix_model.cpp:110-121), plain IEEE intrinsics.
__device__ __noinline__ double epi_residual_ieee(const double *__restrict__ E, double x1, double y1, double x2, double y2)
{
    const double b0 = __dadd_rn(__dadd_rn(__dmul_rn(x2, E[0]), __dmul_rn(y2, E[1])), E[2]);
    const double b1 = __dadd_rn(__dadd_rn(__dmul_rn(x2, E[3]), __dmul_rn(y2, E[4])), E[5]);
    const double b2 = __dadd_rn(__dadd_rn(__dmul_rn(x2, E[6]), __dmul_rn(y2, E[7])), E[8]);
    const double r = __dadd_rn(__dadd_rn(__dmul_rn(b0, x1), __dmul_rn(b1, y1)), b2);
    const double a0 = __dadd_rn(__dadd_rn(__dmul_rn(E[0], x1), __dmul_rn(E[3], y1)), E[6]);
    const double a1 = __dadd_rn(__dadd_rn(__dmul_rn(E[1], x1), __dmul_rn(E[4], y1)), E[7]);
    const double denom = __dadd_rn(
        __dadd_rn(__dadd_rn(__dmul_rn(a0, a0), __dmul_rn(a1, a1)), __dmul_rn(b0, b0)), __dmul_rn(b1, b1));
    if (denom < 1e-20)
        return DBL_MAX;
    return __dsqrt_rn(__ddiv_rn(__dmul_rn(r, r), denom));
}

template <int KIND> __device__ __forceinline__ double residual(const double *__restrict__ M, double x1, double y1, double x2, double y2)
{
    if constexpr (KIND == OCB_MODEL_HOMOGRAPHY)
        return h_residual_ieee(M, x1, y1, x2, y2);
    else
        return epi_residual_ieee(M, x1, y1, x2, y2);
}

// The same residuals as straight-line code (exact_math.cuh): x/z and y/z share one refined reciprocal and nothing
// calls out of line, so the divisions and the square root of one residual overlap in the FP64 pipe.
// Returns e and, through `ratio`, e / thr (given r_thr = rcp_refined(thr)). `rng` (see range_key) comes back
// >= RANGE_OK when an operand left the window in which these forms are equal to div.rn / sqrt.rn; the caller then
// recomputes with residual<KIND>() and __ddiv_rn. An exact fit (e == +0: noise-free scenes) stays in line:
// sqrt(+0) = +0 and quot_shared(+0, thr, r) is +0 / thr exactly.
struct FastResidual
{
    double e, ratio;
    uint32_t rng;
};

__device__ __forceinline__ FastResidual finish_residual(double a, uint32_t rng, double thr, double r_thr)
{
    // a = e^2 >= +0 or NaN. A positive mid_range a passes sqrt_fast_ok and puts e inside the window as well.
    const bool zero = is_pos_zero(a);
    FastResidual f;
    f.e = sqrt_fast(zero ? 1.0 : a);
    f.e = zero ? 0.0 : f.e;
    f.ratio = quot_shared(f.e, thr, r_thr);
    const uint32_t tail = max(range_key(a), range_key(f.ratio));
    f.rng = max(rng, zero ? 0u : tail);
    return f;
}

__device__ __forceinline__ FastResidual h_residual_fast(const double *__restrict__ M, double x1, double y1, double x2,
                                                        double y2, double thr, double r_thr, uint32_t thr_rng)
{
    const double px = __dadd_rn(__dadd_rn(__dmul_rn(M[0], x1), __dmul_rn(M[3], y1)), M[6]);
    const double py = __dadd_rn(__dadd_rn(__dmul_rn(M[1], x1), __dmul_rn(M[4], y1)), M[7]);
    const double pz = __dadd_rn(__dadd_rn(__dmul_rn(M[2], x1), __dmul_rn(M[5], y1)), M[8]);
    const double qx = __dadd_rn(__dadd_rn(__dmul_rn(M[9], x2), __dmul_rn(M[12], y2)), M[15]);
    const double qy = __dadd_rn(__dadd_rn(__dmul_rn(M[10], x2), __dmul_rn(M[13], y2)), M[16]);
    const double qz = __dadd_rn(__dadd_rn(__dmul_rn(M[11], x2), __dmul_rn(M[14], y2)), M[17]);
    const double rp = rcp_refined(pz), rq = rcp_refined(qz);
    const double ax = quot_shared(px, pz, rp), ay = quot_shared(py, pz, rp);
    const double bx = quot_shared(qx, qz, rq), by = quot_shared(qy, qz, rq);
    const double dx = __dsub_rn(ax, x2), dy = __dsub_rn(ay, y2);
    const double ex = __dsub_rn(bx, x1), ey = __dsub_rn(by, y1);
    const double fwd = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
    const double bwd = __dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey));
    const double a = __dmul_rn(__dadd_rn(fwd, bwd), 0.5);
    const uint32_t rng = max(max(max(thr_rng, range_key(pz)), max(range_key(qz), range_key(ax))),
                             max(max(range_key(ay), range_key(bx)), range_key(by)));
    return finish_residual(a, rng, thr, r_thr);
}

__device__ __forceinline__ FastResidual epi_residual_fast(const double *__restrict__ E, double x1, double y1, double x2,
                                                          double y2, double thr, double r_thr, uint32_t thr_rng)
{
    const double b0 = __dadd_rn(__dadd_rn(__dmul_rn(x2, E[0]), __dmul_rn(y2, E[1])), E[2]);
    const double b1 = __dadd_rn(__dadd_rn(__dmul_rn(x2, E[3]), __dmul_rn(y2, E[4])), E[5]);
    const double b2 = __dadd_rn(__dadd_rn(__dmul_rn(x2, E[6]), __dmul_rn(y2, E[7])), E[8]);
    const double r = __dadd_rn(__dadd_rn(__dmul_rn(b0, x1), __dmul_rn(b1, y1)), b2);
    const double a0 = __dadd_rn(__dadd_rn(__dmul_rn(E[0], x1), __dmul_rn(E[3], y1)), E[6]);
    const double a1 = __dadd_rn(__dadd_rn(__dmul_rn(E[1], x1), __dmul_rn(E[4], y1)), E[7]);
    const double denom = __dadd_rn(
        __dadd_rn(__dadd_rn(__dmul_rn(a0, a0), __dmul_rn(a1, a1)), __dmul_rn(b0, b0)), __dmul_rn(b1, b1));
    const double rr = __dmul_rn(r, r);
    // q = e^2. r == 0 exactly gives q = +0 (denom inside the window). denom < 1e-20 (DBL_MAX in the reference) is
    // inside the window (1e-20 ~ 2^-66), so it is flagged explicitly and decided by the out-of-line form.
    const double q = quot_shared(rr, denom, rcp_refined(denom));
    const uint32_t rng = max(max(thr_rng, range_key(denom)), denom < 1e-20 ? RANGE_OK : 0u);
    return finish_residual(q, rng, thr, r_thr);
}

template <int KIND>
__device__ __forceinline__ FastResidual residual_fast(const double *__restrict__ M, double x1, double y1, double x2,
                                                      double y2, double thr, double r_thr, uint32_t thr_rng)
{
    if constexpr (KIND == OCB_MODEL_HOMOGRAPHY)
        return h_residual_fast(M, x1, y1, x2, y2, thr, r_thr, thr_rng);
    else
        return epi_residual_fast(M, x1, y1, x2, y2, thr, r_thr, thr_rng);
}

// measurement / measurement.z for both views; NaN when z/z != 1 (z zero, infinite or NaN), which is what the
// reference's third component is in that case and what then poisons every term of its residual.
__device__ __forceinline__ double4 normalise_corr(const double *__restrict__ c7)
{
    const double z1 = c7[2], z2 = c7[5];
    double4 r;
    r.x = __ddiv_rn(c7[0], z1);
    r.y = __ddiv_rn(c7[1], z1);
    r.z = __ddiv_rn(c7[3], z2);
    r.w = __ddiv_rn(c7[4], z2);
    if (!(__ddiv_rn(z1, z1) == 1.0))
        r.x = r.y = __longlong_as_double(0x7ff8000000000000ll);
    if (!(__ddiv_rn(z2, z2) == 1.0))
        r.z = r.w = __longlong_as_double(0x7ff8000000000000ll);
    return r;
}

__global__ void __launch_bounds__(256)
    k2_prepare_kernel(const double *__restrict__ corr7, const uint32_t *__restrict__ order, uint32_t n,
                      double4 *__restrict__ corr4, uint32_t *__restrict__ pos)
{
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n)
        return;
    const uint32_t idx = order ? order[p] : p;
    corr4[p] = normalise_corr(corr7 + (size_t)idx * 7);
    if (pos)
        pos[p] = idx;
}

struct K2Shared
{
    double M[K2_HG][18];
    double contrib[2][K2_HG][K2_TP + 2]; // +2: the 8 rows the sum warp reads side by side fall in different banks
    uint32_t mask[2][K2_HG][K2_CW + 1];
    alignas(8) uint64_t full_bar[2];  // slab b holds a complete round (K2_CW arrivals, one per compute warp)
    alignas(8) uint64_t empty_bar[2]; // the sum warp is done with slab b (1 arrival)
};

// The scoring loop of one CTA: hypotheses [h0, h0 + nh) of `models` against n evaluation positions.
// Correspondences come either prepared (corr4: [n] pre-divided, already in evaluation order) or raw (c7: [n][7] rows as
// given, read through `order` when it is set and normalised on the fly). W = hypotheses in flight per thread.
// Compute warps and the sum warp meet only through the two mbarrier pairs of the double-buffered slab (no CTA-wide
// barrier per round): a compute warp may run up to two rounds ahead of the slowest one. Ordering: a warp's slab
// accesses -> __syncwarp -> lane 0's mbarrier.arrive (release.cta) -> the other side's try_wait.parity (acquire.cta).
// A barrier is never more than one phase ahead of a waiter, so the parity test is unambiguous. (compute-sanitizer's
// racecheck does not follow try_wait.parity phases and lists the slab accesses as hazards; memcheck and synccheck are
// clean and the scores are compared bit for bit against the oracle in every test.)
// SUBS > 1: the CTA holds SUBS such groups side by side (256 threads each, `sm` is this thread's group); nh == 0 marks a
// group without work. The groups of one CTA meet at a CTA-wide barrier every K2_LOCKSTEP rounds so that none of them
// runs ahead: four independent CTAs per SM drift apart (the warp scheduler favours the oldest), the early ones retire
// and the SM finishes the launch with a quarter of its warps.
template <int KIND, int W, int SUBS>
__device__ __forceinline__ void score_group(K2Shared &sm, const double *__restrict__ models, uint32_t h0, uint32_t nh,
                                            const double4 *__restrict__ corr4, const double *__restrict__ c7,
                                            const uint32_t *__restrict__ order, uint32_t n, double thr,
                                            double *__restrict__ score, uint32_t *__restrict__ count,
                                            uint32_t *__restrict__ bits_pos, uint32_t words)
{
    const uint32_t tid = threadIdx.x % K2_THREADS, warp = tid >> 5, lane = tid & 31;
    for (uint32_t i = tid; i < nh * 18; i += K2_THREADS)
        sm.M[i / 18][i % 18] = models[(size_t)h0 * 18 + i];
    if (tid == 0)
    {
#pragma unroll
        for (int b = 0; b < 2; b++)
        {
            mbar_init(&sm.full_bar[b], K2_CW);
            mbar_init(&sm.empty_bar[b], 1);
        }
        mbar_fence_init();
    }
    __syncthreads();

    const uint32_t rounds = (n + K2_TP - 1) / K2_TP;
    // One loop over the rounds for both roles (and for groups without work), so that the lock-step barrier is ONE
    // __syncthreads reached by every thread of the CTA under a CTA-uniform condition.
    const bool active = nh > 0;
    const bool computes = warp < K2_CW;
    auto load_corr = [&](uint32_t r) {
        const uint32_t p = r * K2_TP + warp * 32 + lane;
        const uint32_t pc = p < n ? p : n - 1; // lanes past the end recompute the last position (never inliers)
        if (corr4)
            return corr4[pc];
        return normalise_corr(c7 + (size_t)(order ? order[pc] : pc) * 7);
    };
    double r_thr = 0.0;     // compute warps: reciprocal shared by every MSAC contribution of this thread
    uint32_t thr_rng = 0;
    double4 c_next = make_double4(0, 0, 0, 0);
    if (active && computes)
    {
        r_thr = rcp_refined(thr);
        thr_rng = range_key(thr);
        if (rounds)
            c_next = load_corr(0);
    }
    double s = 0.0; // sum warp: lane g accumulates hypothesis g
    uint32_t cnt = 0;
    for (uint32_t r = 0; r < rounds; r++)
    {
        if (SUBS > 1 && r > 0 && r % K2_LOCKSTEP == 0)
            __syncthreads();
        if (!active)
            continue;
        if (computes)
        {
            const uint32_t b = r & 1;
            const uint32_t p = r * K2_TP + warp * 32 + lane;
            const bool valid = p < n;
            const double4 c = c_next;
            if (r + 1 < rounds)
                c_next = load_corr(r + 1); // in flight while this round computes
            if (r >= 2)
                mbar_wait(&sm.empty_bar[b], ((r >> 1) - 1) & 1); // the sum warp has consumed round r - 2
            for (uint32_t g = 0; g < nh; g += W)
            {
                FastResidual f[W];
#pragma unroll
                for (int w = 0; w < W; w++)
                    f[w] = residual_fast<KIND>(sm.M[min(g + w, nh - 1)], c.x, c.y, c.z, c.w, thr, r_thr, thr_rng);
#pragma unroll
                for (int w = 0; w < W; w++)
                {
                    if (f[w].rng >= RANGE_OK) // rare: tiny, huge or non-finite operands
                    {
                        f[w].e = residual<KIND>(sm.M[min(g + w, nh - 1)], c.x, c.y, c.z, c.w);
                        f[w].ratio = __ddiv_rn(f[w].e, thr);
                    }
                }
#pragma unroll
                for (int w = 0; w < W; w++)
                {
                    if (g + w < nh)
                    {
                        const bool inl = valid && (f[w].e < thr); // strict, ransac.cpp:189
                        // an outlier parks +0.0: score + 0.0 == score bit for bit (the score is never -0.0), so the
                        // sum warp adds whole 32-position words in order without testing single bits
                        sm.contrib[b][g + w][warp * 32 + lane] =
                            inl ? __dsub_rn(1.0, __dmul_rn(f[w].ratio, f[w].ratio)) : 0.0;
                        const uint32_t m = __ballot_sync(0xFFFFFFFFu, inl);
                        sm.mask[b][g + w][warp] = m; // every lane stores the same word
                        if (bits_pos && lane == 0 && (p >> 5) < words)
                            bits_pos[(size_t)(h0 + g + w) * words + (p >> 5)] = m;
                    }
                }
            }
            __syncwarp();
            if (lane == 0)
                mbar_arrive(&sm.full_bar[r & 1]);
        }
        else
        {
            const uint32_t b = r & 1;
            mbar_wait(&sm.full_bar[b], (r >> 1) & 1);
            if (lane < nh)
            {
#pragma unroll 1
                for (int w = 0; w < K2_CW; w++)
                {
                    const uint32_t m = sm.mask[b][lane][w];
                    if (m == 0)
                        continue; // no inlier among these 32 positions: nothing to add
                    cnt += __popc(m);
                    const double2 *row = reinterpret_cast<const double2 *>(&sm.contrib[b][lane][w * 32]);
#pragma unroll
                    for (int i = 0; i < 16; i++)
                    {
                        const double2 v = row[i];
                        s = __dadd_rn(__dadd_rn(s, v.x), v.y); // score += 1.0 - ratio*ratio, in evaluation order
                    }
                }
            }
            __syncwarp();
            if (lane == 0)
                mbar_arrive(&sm.empty_bar[r & 1]);
        }
    }
    if (active && !computes && lane < nh)
    {
        score[h0 + lane] = s;
        count[h0 + lane] = cnt;
    }
}

template <int KIND, int W, int SUBS>
__global__ void __launch_bounds__(K2_THREADS *SUBS, SUBS > 1 ? 1 : (W == 1 ? 4 : 3))
    k2_score_kernel(const double *__restrict__ models, uint32_t h, uint32_t hg, const double4 *__restrict__ corr4, uint32_t n,
                    double thr, double *__restrict__ score, uint32_t *__restrict__ count,
                    uint32_t *__restrict__ bits_pos, uint32_t words)
{
    extern __shared__ __align__(16) unsigned char k2_dynamic_smem[]; // SUBS x K2Shared (above the 48 KB static limit)
    K2Shared *sm = reinterpret_cast<K2Shared *>(k2_dynamic_smem);
    const uint32_t sub = threadIdx.x / K2_THREADS;
    const uint32_t h0 = min((blockIdx.x * SUBS + sub) * hg, h);
    score_group<KIND, W, SUBS>(sm[sub], models, h0, min(hg, h - h0), corr4, nullptr, nullptr, n, thr, score, count,
                               bits_pos, words);
}

// evaluation-position bit masks -> correspondence-index bit masks (only needed when an order is given)
__global__ void __launch_bounds__(256)
    k2_unpermute_bits_kernel(const uint32_t *__restrict__ bits_pos, const uint32_t *__restrict__ pos, uint32_t h,
                             uint32_t n, uint32_t words, uint32_t *__restrict__ bits)
{
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (uint64_t)h * words)
        return;
    const uint32_t hi = (uint32_t)(g / words), w = (uint32_t)(g % words);
    uint32_t m = bits_pos[g];
    while (m)
    {
        const int bit = __ffs(m) - 1;
        m &= m - 1;
        const uint32_t p = w * 32 + bit;
        if (p < n)
        {
            const uint32_t idx = pos[p];
            atomicOr(&bits[(size_t)hi * words + (idx >> 5)], 1u << (idx & 31));
        }
    }
}

template <int KIND>
__global__ void __launch_bounds__(256)
    k2_residuals_kernel(const double *__restrict__ model18, const double *__restrict__ corr7, uint32_t n,
                        double *__restrict__ e)
{
    __shared__ double M[18];
    if (threadIdx.x < 18)
        M[threadIdx.x] = model18[threadIdx.x];
    __syncthreads();
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const double4 c = normalise_corr(corr7 + (size_t)i * 7);
    e[i] = residual<KIND>(M, c.x, c.y, c.z, c.w);
}

// ----------------------------------------------------------------------------------------------------------
// Request-table form: ONE launch serves many small scoring problems (one RANSAC round of a whole batch of image
// pairs: every pair contributes a batch of hypotheses, a single model to evaluate, or a residual fetch). A CTA finds
// its request by binary search over cta_begin and then does exactly what the single-problem kernels do; the
// correspondences are normalised on the fly from the [n][7] rows (through the evaluation order for mode 0), which
// costs 4 divisions per position per CTA and saves the per-problem prepare launches.
// ----------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(K2_THREADS, 4) k2_requests_kernel(const K2Request *__restrict__ requests, uint32_t n_requests)
{
    __shared__ K2Shared sm;
    uint32_t lo = 0, hi = n_requests - 1;
    while (lo < hi)
    {
        const uint32_t mid = (lo + hi + 1) >> 1;
        if (requests[mid].cta_begin <= blockIdx.x)
            lo = mid;
        else
            hi = mid - 1;
    }
    const K2Request rq = requests[lo];
    const uint32_t local = blockIdx.x - rq.cta_begin;
    if (rq.mode == 2)
    {
        // residual fetch: Model::error of one model for 256 correspondences, index order
        if (threadIdx.x < 18)
            sm.M[0][threadIdx.x] = rq.models[threadIdx.x];
        __syncthreads();
        const uint32_t i = local * K2_THREADS + threadIdx.x;
        if (i < rq.n)
        {
            const double4 c = normalise_corr(rq.c7 + (size_t)i * 7);
            rq.e[i] = rq.kind == OCB_MODEL_HOMOGRAPHY ? residual<OCB_MODEL_HOMOGRAPHY>(sm.M[0], c.x, c.y, c.z, c.w)
                                                      : residual<OCB_MODEL_ESSENTIAL>(sm.M[0], c.x, c.y, c.z, c.w);
        }
        return;
    }
    const uint32_t h0 = local * K2_HG;
    const uint32_t nh = min((uint32_t)K2_HG, rq.h - h0);
    if (rq.kind == OCB_MODEL_HOMOGRAPHY)
        score_group<OCB_MODEL_HOMOGRAPHY, 1, 1>(sm, rq.models, h0, nh, nullptr, rq.c7, rq.order, rq.n, rq.thr, rq.score,
                                             rq.count, rq.bits, rq.words);
    else
        score_group<OCB_MODEL_ESSENTIAL, 1, 1>(sm, rq.models, h0, nh, nullptr, rq.c7, rq.order, rq.n, rq.thr, rq.score,
                                            rq.count, rq.bits, rq.words);
}

uint32_t k2_request_ctas(const K2Request &rq)
{
    if (rq.mode == 2)
        return (rq.n + K2_THREADS - 1) / K2_THREADS;
    return (rq.h + K2_HG - 1) / K2_HG;
}

int k2_run_requests(const K2Request *d_requests, size_t n_requests, uint32_t total_ctas, cudaStream_t stream)
{
    if (n_requests == 0 || total_ctas == 0)
        return 0;
    k2_requests_kernel<<<total_ctas, K2_THREADS, 0, stream>>>(d_requests, (uint32_t)n_requests);
    count_launch();
    OCB_CUDA(cudaGetLastError());
    return 0;
}

int k2_prepare(const double *d_corr7, const uint32_t *d_order, size_t n, double *d_corr4, uint32_t *d_pos,
               cudaStream_t stream)
{
    if (n == 0)
        return 0;
    k2_prepare_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(d_corr7, d_order, (uint32_t)n,
                                                                       reinterpret_cast<double4 *>(d_corr4), d_pos);
    count_launch();
    OCB_CUDA(cudaGetLastError());
    return 0;
}

// Hypotheses per CTA for one launch. A CTA's time grows with its hypotheses and an SM's with the hypotheses of all its
// CTAs, so the launch ends with the most loaded SM: pick the group size in 4..K2_HG that minimises that load
// (4096 hypotheses on 148 SMs: groups of 7 put 28 on every SM, groups of 8 put 32 on 68 SMs and 24 on the rest).
uint32_t k2_pick_group(size_t h, int sms, int resident)
{
    uint32_t best_hg = K2_HG;
    uint64_t best_load = ~0ull;
    for (uint32_t hg = K2_HG; hg >= 4; hg--)
    {
        const uint64_t ctas = (h + hg - 1) / hg;
        const uint64_t per_sm = (ctas + sms - 1) / sms;
        const uint64_t waves = (per_sm + resident - 1) / resident;
        const uint64_t load = (waves > 1 ? waves * resident : per_sm) * hg;
        if (load < best_load)
            best_load = load, best_hg = hg;
    }
    return best_hg;
}

int k2_score(int kind, const double *d_models, size_t h, const double *d_corr4, const uint32_t *d_pos, size_t n,
             double thr, double *d_score, uint32_t *d_count, uint32_t *d_bits, uint32_t *d_bits_scratch,
             cudaStream_t stream)
{
    if (h == 0)
        return 0;
    const uint32_t words = (uint32_t)((n + 31) / 32);
    // with an evaluation order the kernel's ballots are in position order: write them to scratch, then scatter
    uint32_t *bits_pos = d_bits ? (d_pos ? d_bits_scratch : d_bits) : nullptr;
    const double4 *c4 = reinterpret_cast<const double4 *>(d_corr4);
    int dev = 0;
    OCB_CUDA(cudaGetDevice(&dev));
    const int sms = sm_count(dev);
    // four groups are resident per SM either way (64 registers x 256 threads each)
    const uint32_t hg = options().k2_hg >= 1 && options().k2_hg <= K2_HG ? (uint32_t)options().k2_hg
                                                                          : k2_pick_group(h, sms, K2_SUBS);
    const uint32_t groups = (uint32_t)((h + hg - 1) / hg);
    // lock-step form (K2_SUBS groups in one 1024-thread CTA) once every SM would host several groups anyway;
    // option k2_variant: 0 = by size, 1 = always independent CTAs, 2 = always lock-step
    const int v = options().k2_variant;
    const bool lockstep = v == 2 || (v == 0 && groups > 2u * (uint32_t)sms);
    static std::atomic<bool> attr_set_on[64]; // the attribute is per device; setting it twice is harmless
    std::atomic<bool> &attr_set = attr_set_on[dev & 63];
    if (lockstep && !attr_set.load(std::memory_order_acquire))
    {
        const int bytes = (int)(K2_SUBS * sizeof(K2Shared));
        OCB_CUDA(cudaFuncSetAttribute(k2_score_kernel<OCB_MODEL_HOMOGRAPHY, 1, K2_SUBS>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        OCB_CUDA(cudaFuncSetAttribute(k2_score_kernel<OCB_MODEL_ESSENTIAL, 1, K2_SUBS>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        attr_set.store(true, std::memory_order_release);
    }
#define OCB_K2_LAUNCH(KIND, SUBS)                                                                                      \
    k2_score_kernel<KIND, 1, SUBS><<<(groups + SUBS - 1) / SUBS, K2_THREADS * SUBS, SUBS * sizeof(K2Shared), stream>>>( \
        d_models, (uint32_t)h, hg, c4, (uint32_t)n, thr, d_score, d_count, bits_pos, words)
    if (kind == OCB_MODEL_HOMOGRAPHY)
    {
        if (lockstep)
            OCB_K2_LAUNCH(OCB_MODEL_HOMOGRAPHY, K2_SUBS);
        else
            OCB_K2_LAUNCH(OCB_MODEL_HOMOGRAPHY, 1);
    }
    else
    {
        if (lockstep)
            OCB_K2_LAUNCH(OCB_MODEL_ESSENTIAL, K2_SUBS);
        else
            OCB_K2_LAUNCH(OCB_MODEL_ESSENTIAL, 1);
    }
#undef OCB_K2_LAUNCH
    count_launch();
    OCB_CUDA(cudaGetLastError());
    if (d_bits && d_pos && words > 0)
    {
        OCB_CUDA(cudaMemsetAsync(d_bits, 0, (size_t)h * words * sizeof(uint32_t), stream));
        const uint64_t total = (uint64_t)h * words;
        k2_unpermute_bits_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(bits_pos, d_pos, (uint32_t)h,
                                                                                    (uint32_t)n, words, d_bits);
        count_launch();
        OCB_CUDA(cudaGetLastError());
    }
    return 0;
}

int k2_residuals(int kind, const double *d_model18, const double *d_corr7, size_t n, double *d_e, cudaStream_t stream)
{
    if (n == 0)
        return 0;
    const unsigned grid = (unsigned)((n + 255) / 256);
    if (kind == OCB_MODEL_HOMOGRAPHY)
        k2_residuals_kernel<OCB_MODEL_HOMOGRAPHY><<<grid, 256, 0, stream>>>(d_model18, d_corr7, (uint32_t)n, d_e);
    else
        k2_residuals_kernel<OCB_MODEL_ESSENTIAL><<<grid, 256, 0, stream>>>(d_model18, d_corr7, (uint32_t)n, d_e);
    count_launch();
    OCB_CUDA(cudaGetLastError());
    return 0;
}

} // namespace ocb
