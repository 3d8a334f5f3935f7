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
// double-buffered shared-memory slab. Warp 7 ("sum") runs one round behind: lane g walks hypothesis g's
// masks and adds the parked contributions of the inliers in order. Adding nothing for an outlier is
// exactly what the reference does, so only set bits are visited.
#include "ocb_internal.cuh"

#include <cfloat>

namespace ocb
{

constexpr int K2_HG = 8;          // hypotheses per CTA
constexpr int K2_CW = 7;          // compute warps
constexpr int K2_TP = K2_CW * 32; // evaluation positions per round
constexpr int K2_THREADS = (K2_CW + 1) * 32;

// homography_model::error (homography_model.cpp:89-97) on pre-divided coordinates.
// M = [H (9, column-major) | H^-1 (9)].
__device__ __forceinline__ double h_residual(const double *__restrict__ M, double x1, double y1, double x2, double y2)
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
// (essential_matrix_model.cpp:112-123, fundamental_matrix_model.cpp:110-121).
__device__ __forceinline__ double epi_residual(const double *__restrict__ E, double x1, double y1, double x2, double y2)
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
        return h_residual(M, x1, y1, x2, y2);
    else
        return epi_residual(M, x1, y1, x2, y2);
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

template <int KIND>
__global__ void __launch_bounds__(K2_THREADS)
    k2_score_kernel(const double *__restrict__ models, uint32_t h, const double4 *__restrict__ corr4, uint32_t n,
                    double thr, double *__restrict__ score, uint32_t *__restrict__ count,
                    uint32_t *__restrict__ bits_pos, uint32_t words)
{
    __shared__ double M[K2_HG][18];
    __shared__ double contrib[2][K2_HG][K2_TP];
    __shared__ uint32_t mask[2][K2_HG][K2_CW + 1];

    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t h0 = blockIdx.x * K2_HG;
    const uint32_t nh = min((uint32_t)K2_HG, h - h0);
    for (uint32_t i = tid; i < nh * 18; i += K2_THREADS)
        M[i / 18][i % 18] = models[(size_t)h0 * 18 + i];
    __syncthreads();

    const uint32_t rounds = (n + K2_TP - 1) / K2_TP;
    double s = 0.0;
    uint32_t cnt = 0;
    for (uint32_t r = 0; r <= rounds; r++)
    {
        if (warp < K2_CW)
        {
            if (r < rounds)
            {
                const uint32_t p = r * K2_TP + warp * 32 + lane;
                const bool valid = p < n;
                double4 c = make_double4(0, 0, 0, 0);
                if (valid)
                    c = corr4[p];
#pragma unroll
                for (int g = 0; g < K2_HG; g++)
                {
                    if ((uint32_t)g < nh)
                    {
                        const double e = residual<KIND>(M[g], c.x, c.y, c.z, c.w);
                        const bool inl = valid && (e < thr); // strict, ransac.cpp:189
                        const double ratio = __ddiv_rn(e, thr);
                        contrib[r & 1][g][warp * 32 + lane] = __dsub_rn(1.0, __dmul_rn(ratio, ratio));
                        const uint32_t m = __ballot_sync(0xFFFFFFFFu, inl);
                        if (lane == 0)
                        {
                            mask[r & 1][g][warp] = m;
                            if (bits_pos && (p >> 5) < words)
                                bits_pos[(size_t)(h0 + g) * words + (p >> 5)] = m;
                        }
                    }
                }
            }
        }
        else if (r > 0 && lane < nh)
        {
            const uint32_t b = (r - 1) & 1;
#pragma unroll
            for (int w = 0; w < K2_CW; w++)
            {
                uint32_t m = mask[b][lane][w];
                cnt += __popc(m);
                while (m)
                {
                    const int bit = __ffs(m) - 1;
                    s = __dadd_rn(s, contrib[b][lane][w * 32 + bit]); // score += 1.0 - ratio*ratio, in order
                    m &= m - 1;
                }
            }
        }
        __syncthreads();
    }
    if (warp == K2_CW && lane < nh)
    {
        score[h0 + lane] = s;
        count[h0 + lane] = cnt;
    }
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
template <int KIND>
__device__ __forceinline__ void score_request(const K2Request &rq, uint32_t local_cta, double (*M)[18],
                                              double (*contrib)[K2_HG][K2_TP], uint32_t (*mask)[K2_HG][K2_CW + 1])
{
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t h0 = local_cta * K2_HG;
    const uint32_t nh = min((uint32_t)K2_HG, rq.h - h0);
    const uint32_t n = rq.n;
    const double thr = rq.thr;
    for (uint32_t i = tid; i < nh * 18; i += K2_THREADS)
        M[i / 18][i % 18] = rq.models[(size_t)h0 * 18 + i];
    __syncthreads();
    const uint32_t rounds = (n + K2_TP - 1) / K2_TP;
    double s = 0.0;
    uint32_t cnt = 0;
    for (uint32_t r = 0; r <= rounds; r++)
    {
        if (warp < K2_CW)
        {
            if (r < rounds)
            {
                const uint32_t p = r * K2_TP + warp * 32 + lane;
                const bool valid = p < n;
                double4 c = make_double4(0, 0, 0, 0);
                if (valid)
                {
                    const uint32_t idx = rq.order ? rq.order[p] : p;
                    c = normalise_corr(rq.c7 + (size_t)idx * 7);
                }
#pragma unroll
                for (int g = 0; g < K2_HG; g++)
                {
                    if ((uint32_t)g < nh)
                    {
                        const double e = residual<KIND>(M[g], c.x, c.y, c.z, c.w);
                        const bool inl = valid && (e < thr);
                        const double ratio = __ddiv_rn(e, thr);
                        contrib[r & 1][g][warp * 32 + lane] = __dsub_rn(1.0, __dmul_rn(ratio, ratio));
                        const uint32_t m = __ballot_sync(0xFFFFFFFFu, inl);
                        if (lane == 0)
                        {
                            mask[r & 1][g][warp] = m;
                            if (rq.bits && (p >> 5) < rq.words)
                                rq.bits[(size_t)(h0 + g) * rq.words + (p >> 5)] = m;
                        }
                    }
                }
            }
        }
        else if (r > 0 && lane < nh)
        {
            const uint32_t b = (r - 1) & 1;
#pragma unroll
            for (int w = 0; w < K2_CW; w++)
            {
                uint32_t m = mask[b][lane][w];
                cnt += __popc(m);
                while (m)
                {
                    const int bit = __ffs(m) - 1;
                    s = __dadd_rn(s, contrib[b][lane][w * 32 + bit]);
                    m &= m - 1;
                }
            }
        }
        __syncthreads();
    }
    if (warp == K2_CW && lane < nh)
    {
        rq.score[h0 + lane] = s;
        rq.count[h0 + lane] = cnt;
    }
}

__global__ void __launch_bounds__(K2_THREADS) k2_requests_kernel(const K2Request *__restrict__ requests, uint32_t n_requests)
{
    __shared__ double M[K2_HG][18];
    __shared__ double contrib[2][K2_HG][K2_TP];
    __shared__ uint32_t mask[2][K2_HG][K2_CW + 1];
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
            M[0][threadIdx.x] = rq.models[threadIdx.x];
        __syncthreads();
        const uint32_t i = local * K2_THREADS + threadIdx.x;
        if (i < rq.n)
        {
            const double4 c = normalise_corr(rq.c7 + (size_t)i * 7);
            rq.e[i] = rq.kind == OCB_MODEL_HOMOGRAPHY ? residual<OCB_MODEL_HOMOGRAPHY>(M[0], c.x, c.y, c.z, c.w)
                                                      : residual<OCB_MODEL_ESSENTIAL>(M[0], c.x, c.y, c.z, c.w);
        }
        return;
    }
    if (rq.kind == OCB_MODEL_HOMOGRAPHY)
        score_request<OCB_MODEL_HOMOGRAPHY>(rq, local, M, contrib, mask);
    else
        score_request<OCB_MODEL_ESSENTIAL>(rq, local, M, contrib, mask);
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

int k2_score(int kind, const double *d_models, size_t h, const double *d_corr4, const uint32_t *d_pos, size_t n,
             double thr, double *d_score, uint32_t *d_count, uint32_t *d_bits, uint32_t *d_bits_scratch,
             cudaStream_t stream)
{
    if (h == 0)
        return 0;
    const uint32_t words = (uint32_t)((n + 31) / 32);
    // with an evaluation order the kernel's ballots are in position order: write them to scratch, then scatter
    uint32_t *bits_pos = d_bits ? (d_pos ? d_bits_scratch : d_bits) : nullptr;
    const unsigned grid = (unsigned)((h + K2_HG - 1) / K2_HG);
    const double4 *c4 = reinterpret_cast<const double4 *>(d_corr4);
    if (kind == OCB_MODEL_HOMOGRAPHY)
        k2_score_kernel<OCB_MODEL_HOMOGRAPHY>
            <<<grid, K2_THREADS, 0, stream>>>(d_models, (uint32_t)h, c4, (uint32_t)n, thr, d_score, d_count, bits_pos, words);
    else
        k2_score_kernel<OCB_MODEL_ESSENTIAL>
            <<<grid, K2_THREADS, 0, stream>>>(d_models, (uint32_t)h, c4, (uint32_t)n, thr, d_score, d_count, bits_pos, words);
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
