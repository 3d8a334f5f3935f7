// K3 -- homography fits on the device (sm_100a), bit for bit the host / oracle arithmetic.
//
// Replaces, for a whole batch of RANSAC hypotheses, homography_model::checkSampleDegeneracy
// (reference src/model_inliers/homography_model.cpp:120-136) and homography_model::fit (:19-50): the 4-point DLT
// system with the h33 == 1 row (:26-41), P.fullPivLu().solve(rhs) (:44), the renormalisation by H(2,2) (:48) and
// homography.inverse() (:49). The all-inlier refit of the local-optimisation loop (fitInliers, :52-87, a
// (2m+1) x 9 system) stays on the host (host/models.cpp): it runs a handful of times per RANSAC run.
//
// Exactness: Eigen is not under /root/reference (un-vendored find_package dependency), so the algorithms are the
// published Eigen 3.4 ones as restated in host/linalg.cpp (FullPivLU with complete pivoting, cofactor inverse for
// 3x3). Every multiply / subtract / divide below goes through the *_rn intrinsics (never contracted into FMAs) in
// the order of host/linalg.cpp, so a device fit equals the host fit and the oracle's fit bit for bit; the tests
// compare them with array_equal.
//
// Mapping
//   * minimal-sample fits (k3_fit_samples_kernel): one THREAD per hypothesis. The 9x9 system lives in shared
//     memory, element-major ([81][32 threads]), so the data-dependent pivot indexing costs no local-memory traffic
//     and no bank conflicts (every thread of the warp touches its own bank column).
#include "ocb_internal.cuh"

#include <cfloat>

namespace ocb
{

constexpr int K3_THREADS = 32;

__device__ __forceinline__ double k3_nan()
{
    return __longlong_as_double(0x7ff8000000000000ll);
}

// std::max(a, b) of libstdc++: (a < b) ? b : a  -- a NaN in b loses, a NaN in a sticks
__device__ __forceinline__ double std_max(double a, double b)
{
    return (a < b) ? b : a;
}

// cofactor inverse of a column-major 3x3 (host/linalg.cpp invert3 == Eigen's compute_inverse_size3)
__device__ __forceinline__ void invert3_dev(const double *m, double *out)
{
#define K3_AT(r, c) m[(r) + 3 * (c)]
#define K3_MINOR(i, j)                                                                                                 \
    __dsub_rn(__dmul_rn(K3_AT(((i) + 1) % 3, ((j) + 1) % 3), K3_AT(((i) + 2) % 3, ((j) + 2) % 3)),                      \
              __dmul_rn(K3_AT(((i) + 1) % 3, ((j) + 2) % 3), K3_AT(((i) + 2) % 3, ((j) + 1) % 3)))
    const double k00 = K3_MINOR(0, 0), k10 = K3_MINOR(1, 0), k20 = K3_MINOR(2, 0);
    const double det =
        __dadd_rn(__dadd_rn(__dmul_rn(k00, K3_AT(0, 0)), __dmul_rn(k10, K3_AT(1, 0))), __dmul_rn(k20, K3_AT(2, 0)));
    const double s = __ddiv_rn(1.0, det);
    out[0] = __dmul_rn(k00, s);
    out[3] = __dmul_rn(k10, s);
    out[6] = __dmul_rn(k20, s);
    out[1] = __dmul_rn(K3_MINOR(0, 1), s), out[4] = __dmul_rn(K3_MINOR(1, 1), s), out[7] = __dmul_rn(K3_MINOR(2, 1), s);
    out[2] = __dmul_rn(K3_MINOR(0, 2), s), out[5] = __dmul_rn(K3_MINOR(1, 2), s), out[8] = __dmul_rn(K3_MINOR(2, 2), s);
#undef K3_MINOR
#undef K3_AT
}

// solution h[9] (row-major homography entries) -> M18 = [H column-major / H(2,2) | inverse]
__device__ __forceinline__ void finish_model(const double *h, double *M18)
{
    double H[9];
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 3; c++)
            H[r + 3 * c] = h[3 * r + c];
    const double h22 = H[8];
#pragma unroll
    for (int i = 0; i < 9; i++)
        H[i] = __ddiv_rn(H[i], h22);
    double G[9];
    invert3_dev(H, G);
#pragma unroll
    for (int i = 0; i < 9; i++)
        M18[i] = H[i], M18[9 + i] = G[i];
}

// ----------------------------------------------------------------------------------------------------------
// minimal-sample fits: one thread per hypothesis
// ----------------------------------------------------------------------------------------------------------
#define A_(r, c) A[(r) + 9 * (c)][lane]

__global__ void __launch_bounds__(K3_THREADS)
    k3_fit_samples_kernel(const K3FitJob *__restrict__ jobs, uint32_t n_jobs, uint32_t total)
{
    __shared__ double A[81][K3_THREADS];
    __shared__ double Y[9][K3_THREADS];
    const uint32_t lane = threadIdx.x;
    const uint32_t t = blockIdx.x * K3_THREADS + lane;
    if (t >= total)
        return;
    uint32_t lo = 0, hi = n_jobs - 1;
    while (lo < hi)
    {
        const uint32_t mid = (lo + hi + 1) >> 1;
        if (jobs[mid].begin <= t)
            lo = mid;
        else
            hi = mid - 1;
    }
    const K3FitJob job = jobs[lo];
    const uint32_t hyp = t - job.begin;
    const uint32_t *smp = job.samples + (size_t)hyp * 4;
    double *out = job.models_out + (size_t)hyp * 18;

    // hnormalized sample points (homography_model.cpp:26-29, :124-127): two true divisions per view
    double x[4], y[4], u[4], v[4];
#pragma unroll
    for (int i = 0; i < 4; i++)
    {
        const double *c7 = job.c7 + (size_t)smp[i] * 7;
        x[i] = __ddiv_rn(c7[0], c7[2]);
        y[i] = __ddiv_rn(c7[1], c7[2]);
        u[i] = __ddiv_rn(c7[3], c7[5]);
        v[i] = __ddiv_rn(c7[4], c7[5]);
    }
    // checkSampleDegeneracy (:120-136): any three source points (nearly) collinear
    bool degenerate = false;
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = a + 1; b < 4; b++)
#pragma unroll
            for (int c = b + 1; c < 4; c++)
            {
                const double ux = __dsub_rn(x[b], x[a]), uy = __dsub_rn(y[b], y[a]);
                const double wx = __dsub_rn(x[c], x[a]), wy = __dsub_rn(y[c], y[a]);
                if (fabs(__dsub_rn(__dmul_rn(ux, wy), __dmul_rn(uy, wx))) < 1e-10)
                    degenerate = true;
            }
    if (job.degenerate)
        job.degenerate[hyp] = degenerate ? 1 : 0;
    if (degenerate)
    {
        // the reference skips the iteration (ransac.cpp:173-177): no model; NaN scores nothing
        for (int i = 0; i < 18; i++)
            out[i] = k3_nan();
        return;
    }

    // P (9x9, column-major) and rhs = e8 (:26-41)
#pragma unroll
    for (int i = 0; i < 4; i++)
    {
        const double top[9] = {-x[i], -y[i], -1.0, 0.0, 0.0, 0.0, __dmul_rn(x[i], u[i]), __dmul_rn(y[i], u[i]), u[i]};
        const double bot[9] = {0.0, 0.0, 0.0, -x[i], -y[i], -1.0, __dmul_rn(x[i], v[i]), __dmul_rn(y[i], v[i]), v[i]};
#pragma unroll
        for (int k = 0; k < 9; k++)
        {
            A_(2 * i, k) = top[k];
            A_(2 * i + 1, k) = bot[k];
        }
    }
#pragma unroll
    for (int k = 0; k < 9; k++)
        A_(8, k) = k == 8 ? 1.0 : 0.0;

    // FullPivLU (host/linalg.cpp full_piv_lu_solve, R = Cn = K = 9)
    unsigned long long row_swap = 0, col_swap = 0; // 4 bits per step
    int pivots = 9;
    double biggest_pivot = 0.0;
    for (int k = 0; k < 9; ++k)
    {
        int pr = k, pc = k;
        double best = fabs(A_(k, k));
        for (int c = k; c < 9; ++c)
        {
            double cmax = 0.0;
            for (int r = k; r < 9; ++r)
                cmax = std_max(cmax, fabs(A_(r, c)));
            if (cmax > best)
            {
                best = cmax, pc = c;
                for (int r = k; r < 9; ++r)
                    if (fabs(A_(r, c)) == cmax)
                    {
                        pr = r;
                        break;
                    }
            }
        }
        if (best == 0.0)
        {
            pivots = k;
            for (int i = k; i < 9; ++i)
                row_swap |= (unsigned long long)i << (4 * i), col_swap |= (unsigned long long)i << (4 * i);
            break;
        }
        biggest_pivot = std_max(biggest_pivot, best);
        row_swap |= (unsigned long long)pr << (4 * k);
        col_swap |= (unsigned long long)pc << (4 * k);
        if (pr != k)
            for (int c = 0; c < 9; ++c)
            {
                const double tmp = A_(k, c);
                A_(k, c) = A_(pr, c);
                A_(pr, c) = tmp;
            }
        if (pc != k)
            for (int r = 0; r < 9; ++r)
            {
                const double tmp = A_(r, k);
                A_(r, k) = A_(r, pc);
                A_(r, pc) = tmp;
            }
        const double d = A_(k, k);
        for (int r = k + 1; r < 9; ++r)
            A_(r, k) = __ddiv_rn(A_(r, k), d);
        for (int c = k + 1; c < 9; ++c)
        {
            const double top = A_(k, c);
            for (int r = k + 1; r < 9; ++r)
                A_(r, c) = __dsub_rn(A_(r, c), __dmul_rn(A_(r, k), top));
        }
    }
    const double cut = __dmul_rn(fabs(biggest_pivot), DBL_EPSILON * 9.0);
    int rank = 0;
    for (int i = 0; i < pivots; ++i)
        if (fabs(A_(i, i)) > cut)
            ++rank;

    double h[9];
#pragma unroll
    for (int i = 0; i < 9; i++)
        h[i] = 0.0;
    if (rank > 0)
    {
#define Y_(i) Y[i][lane]
        for (int i = 0; i < 9; i++)
            Y_(i) = i == 8 ? 1.0 : 0.0;
        for (int k = 0; k < 9; ++k)
        {
            const int p = (int)((row_swap >> (4 * k)) & 15);
            const double tmp = Y_(k);
            Y_(k) = Y_(p);
            Y_(p) = tmp;
        }
        for (int c = 0; c < 9; ++c)
        {
            const double yc = Y_(c);
            if (yc == 0.0)
                continue;
            for (int r = c + 1; r < 9; ++r)
                Y_(r) = __dsub_rn(Y_(r), __dmul_rn(yc, A_(r, c)));
        }
        for (int c = rank - 1; c >= 0; --c)
        {
            if (Y_(c) == 0.0)
                continue;
            Y_(c) = __ddiv_rn(Y_(c), A_(c, c));
            const double yc = Y_(c);
            for (int r = 0; r < c; ++r)
                Y_(r) = __dsub_rn(Y_(r), __dmul_rn(yc, A_(r, c)));
        }
        for (int i = rank; i < 9; ++i)
            Y_(i) = 0.0;
        for (int k = 8; k >= 0; --k)
        {
            const int p = (int)((col_swap >> (4 * k)) & 15);
            const double tmp = Y_(k);
            Y_(k) = Y_(p);
            Y_(p) = tmp;
        }
#pragma unroll
        for (int i = 0; i < 9; i++)
            h[i] = Y_(i);
#undef Y_
    }
    double M18[18];
    finish_model(h, M18);
#pragma unroll
    for (int i = 0; i < 18; i++)
        out[i] = M18[i];
}
#undef A_

int k3_fit_samples(const K3FitJob *d_jobs, size_t n_jobs, uint32_t total, cudaStream_t stream)
{
    if (n_jobs == 0 || total == 0)
        return 0;
    k3_fit_samples_kernel<<<(total + K3_THREADS - 1) / K3_THREADS, K3_THREADS, 0, stream>>>(d_jobs, (uint32_t)n_jobs,
                                                                                           total);
    count_launch();
    OCB_CUDA(cudaGetLastError());
    return 0;
}

} // namespace ocb
