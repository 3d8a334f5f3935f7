// K3 -- homography fits on the device (sm_100a), bit for bit the host / oracle arithmetic.
//
// Replaces, for a whole batch of RANSAC hypotheses, homography_model::checkSampleDegeneracy
// (reference src/model_inliers/homography_model.cpp:120-136) and homography_model::fit (:19-50): the 4-point DLT
// system with the h33 == 1 row (:26-41), P.fullPivLu().solve(rhs) (:44), the renormalisation by H(2,2) (:48) and
// homography.inverse() (:49); and homography_model::fitInliers (:52-87), the all-inlier refit of the local-optimisation
// loop of ransac.cpp:224-245: the same DLT system with two rows per inlier, (2m+1) x 9, one CTA per system.
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
//   * all-inlier refits (k3_fit_inliers_kernel): one CTA per system, the tall matrix ROW-major in global memory
//     (L2 resident: 144 bytes per inlier), every thread owning whole rows with two of them in flight. An elimination
//     step is one pass over the rows (column swap, multiplier, rank-1 update) that also collects the next step's
//     pivot search, then a CTA-wide arg-max with the sequential scan's tie rule (largest magnitude, then smallest
//     column, then smallest row); each element sees exactly the host's operations, so the thread order does not matter.
//     On the host this LU was the hottest routine of the RANSAC tail (0.1 ms per refit, two or three per image pair).
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


// ----------------------------------------------------------------------------------------------------------
// all-inlier refits: one CTA per system
// ----------------------------------------------------------------------------------------------------------
constexpr int K3I_THREADS = 256;

struct PivotKey
{
    double mag;
    int c, r;
};
// true when b has to replace a: the sequential column-by-column scan of host/linalg.cpp keeps the FIRST element of
// the largest magnitude (a strict comparison; NaNs lose every comparison)
__device__ __forceinline__ bool pivot_beats(const PivotKey &b, const PivotKey &a)
{
    return b.mag > a.mag || (b.mag == a.mag && (b.c < a.c || (b.c == a.c && b.r < a.r)));
}

__global__ void __launch_bounds__(K3I_THREADS, 2) k3_fit_inliers_kernel(const K3InlierJob *__restrict__ jobs)
{
    __shared__ uint32_t warp_sum[K3I_THREADS / 32];
    __shared__ PivotKey warp_key[K3I_THREADS / 32];
    __shared__ PivotKey chosen;
    __shared__ uint32_t total_sh;
    __shared__ double rowk[9];
    const K3InlierJob job = jobs[blockIdx.x];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t n = job.n, words = (n + 31) / 32;
    auto word_at = [&](uint32_t w) -> uint32_t {
        uint32_t m = job.bits[w];
        if (w == words - 1 && (n & 31))
            m &= (1u << (n & 31)) - 1u;
        return m;
    };
    // CTA-wide arg-max of the threads' keys -> chosen. `kk` = the diagonal element the sequential scan starts from: when
    // it is NaN nothing replaces it there (every comparison with a NaN is false), so it stays the pivot.
    auto reduce_keys = [&](PivotKey key, const double *kk) {
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1)
        {
            PivotKey o;
            o.mag = __shfl_xor_sync(0xFFFFFFFFu, key.mag, d);
            o.c = __shfl_xor_sync(0xFFFFFFFFu, key.c, d);
            o.r = __shfl_xor_sync(0xFFFFFFFFu, key.r, d);
            if (pivot_beats(o, key))
                key = o;
        }
        if (lane == 0)
            warp_key[warp] = key;
        __syncthreads(); // also: every store of the pass before is visible to the CTA
        if (tid == 0)
        {
            PivotKey best = warp_key[0];
            for (int i = 1; i < K3I_THREADS / 32; i++)
                if (pivot_beats(warp_key[i], best))
                    best = warp_key[i];
            const double diag = fabs(*kk);
            if (diag != diag)
                best.mag = diag;
            chosen = best;
        }
        __syncthreads();
    };

    // ---- m = number of inliers (fitInliers gathers them in index order, homography_model.cpp:56-62)
    uint32_t cnt = 0;
    for (uint32_t w = tid; w < words; w += K3I_THREADS)
        cnt += __popc(word_at(w));
    cnt = __reduce_add_sync(0xFFFFFFFFu, cnt);
    if (lane == 0)
        warp_sum[warp] = cnt;
    __syncthreads();
    if (tid == 0)
    {
        uint32_t t = 0;
        for (int i = 0; i < K3I_THREADS / 32; i++)
            t += warp_sum[i];
        total_sh = t;
    }
    __syncthreads();
    const uint32_t m = total_sh;
    const int R = (int)(2 * m + 1);
    double *__restrict__ P = job.P; // [R][9] ROW-major: a thread owns whole rows, so an elimination step is one pass
#define P_(r, c) P[(size_t)(r) * 9 + (c)]

    // ---- rows 2i, 2i+1 of the i-th inlier (:64-73); last row: h33 == 1 (:75-77). The pivot search of the first
    // elimination step rides along.
    PivotKey key{-1.0, 0, 0}; // loses against every element (magnitudes are >= 0; NaNs lose every comparison)
    uint32_t running = 0;
    for (uint32_t base = 0; base < words; base += K3I_THREADS)
    {
        const uint32_t w = base + tid;
        uint32_t mask = w < words ? word_at(w) : 0u;
        const uint32_t c = __popc(mask);
        uint32_t incl = c; // inclusive prefix of c over the warp
#pragma unroll
        for (int d = 1; d < 32; d <<= 1)
        {
            const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= (uint32_t)d)
                incl += v;
        }
        __syncthreads(); // warp_sum of the previous chunk has been read
        if (lane == 31)
            warp_sum[warp] = incl;
        __syncthreads();
        uint32_t before = running + incl - c;
        uint32_t chunk_total = 0;
        for (uint32_t i = 0; i < K3I_THREADS / 32; i++)
        {
            if (i < warp)
                before += warp_sum[i];
            chunk_total += warp_sum[i];
        }
        running += chunk_total;
        uint32_t rank = before;
        while (mask)
        {
            const uint32_t bit = __ffs(mask) - 1;
            mask &= mask - 1;
            const double *c7 = job.c7 + (size_t)(w * 32 + bit) * 7;
            const double x = __ddiv_rn(c7[0], c7[2]), y = __ddiv_rn(c7[1], c7[2]);
            const double u = __ddiv_rn(c7[3], c7[5]), v = __ddiv_rn(c7[4], c7[5]);
            const double top[9] = {-x, -y, -1.0, 0.0, 0.0, 0.0, __dmul_rn(x, u), __dmul_rn(y, u), u};
            const double bot[9] = {0.0, 0.0, 0.0, -x, -y, -1.0, __dmul_rn(x, v), __dmul_rn(y, v), v};
            const int row = (int)(2 * rank);
#pragma unroll
            for (int k = 0; k < 9; k++)
            {
                P_(row, k) = top[k];
                P_(row + 1, k) = bot[k];
                const PivotKey e0{fabs(top[k]), k, row}, e1{fabs(bot[k]), k, row + 1};
                if (pivot_beats(e0, key))
                    key = e0;
                if (pivot_beats(e1, key))
                    key = e1;
            }
            rank++;
        }
    }
    if (tid < 9)
    {
        P_(R - 1, tid) = tid == 8 ? 1.0 : 0.0;
        const PivotKey e{tid == 8 ? 1.0 : 0.0, (int)tid, R - 1};
        if (pivot_beats(e, key))
            key = e;
    }
    __syncthreads();
    reduce_keys(key, &P_(0, 0));

    // ---- FullPivLU (host/linalg.cpp full_piv_lu_solve, Cn = 9, K = min(R, 9)). Step k: swap rows k <-> pr, swap
    // columns k <-> pc, divide column k below the diagonal by the pivot, subtract the rank-1 product from the trailing
    // block - one pass over the rows below k, which also collects the pivot search of step k + 1.
    const int K = R < 9 ? R : 9;
    int row_swap[9], col_swap[9];
    int pivots = K;
    double biggest_pivot = 0.0;
    for (int k = 0; k < K; ++k)
    {
        const double best = chosen.mag;
        const int pr = chosen.r, pc = chosen.c;
        if (best == 0.0)
        {
            pivots = k;
            for (int i = k; i < K; ++i)
                row_swap[i] = col_swap[i] = i;
            break;
        }
        biggest_pivot = std_max(biggest_pivot, best);
        row_swap[k] = pr;
        col_swap[k] = pc;
        __syncthreads(); // everybody has read `chosen`
        if (pr != k && tid < 9)
        {
            const double tmp = P_(k, tid);
            P_(k, tid) = P_(pr, tid);
            P_(pr, tid) = tmp;
        }
        __syncthreads();
        if (tid < 9) // the pivot row, columns k <-> pc swapped
            rowk[tid] = P_(k, tid == (uint32_t)k ? pc : (tid == (uint32_t)pc ? k : (int)tid));
        else if (tid - 9 < (uint32_t)k && pc != k) // the finished rows above it: the column swap only
        {
            const int r = (int)tid - 9;
            const double tmp = P_(r, k);
            P_(r, k) = P_(r, pc);
            P_(r, pc) = tmp;
        }
        __syncthreads();
        if (tid < 9)
            P_(k, tid) = rowk[tid];
        const double d = rowk[k];
        double top[9];
#pragma unroll
        for (int c = 0; c < 9; c++)
            top[c] = rowk[c];
        key = PivotKey{-1.0, 0, 0};
        constexpr int BATCH = 2; // rows in flight per thread (independent loads)
        for (int r0 = k + 1 + (int)tid; r0 < R; r0 += BATCH * K3I_THREADS)
        {
            double v[BATCH][9];
#pragma unroll
            for (int j = 0; j < BATCH; j++)
            {
                const int r = r0 + j * K3I_THREADS;
                if (r < R)
                {
#pragma unroll
                    for (int c = 0; c < 9; c++)
                        if (c >= k)
                            v[j][c] = P_(r, c);
                }
            }
#pragma unroll
            for (int j = 0; j < BATCH; j++)
            {
                const int r = r0 + j * K3I_THREADS;
                if (r < R)
                {
                    // column swap k <-> pc (pc >= k), then the multiplier and the update
                    double vk = 0.0;
#pragma unroll
                    for (int c = 0; c < 9; c++)
                        if (c == k)
                            vk = v[j][c];
#pragma unroll
                    for (int c = 0; c < 9; c++)
                        if (c == pc && c != k)
                        {
                            const double tmp = v[j][c];
                            v[j][c] = vk;
                            vk = tmp;
                        }
                    vk = __ddiv_rn(vk, d);
                    P_(r, k) = vk;
#pragma unroll
                    for (int c = 0; c < 9; c++)
                        if (c > k)
                        {
                            const double nv = __dsub_rn(v[j][c], __dmul_rn(vk, top[c]));
                            P_(r, c) = nv;
                            const PivotKey e{fabs(nv), c, r};
                            if (pivot_beats(e, key))
                                key = e;
                        }
                }
            }
        }
        if (k + 1 < K)
            reduce_keys(key, &P_(k + 1, k + 1));
    }
    __syncthreads();
    if (tid != 0)
        return;

    // ---- the 9 x 9 solve is sequential and tiny: one thread
    const double cut = __dmul_rn(fabs(biggest_pivot), DBL_EPSILON * (double)K);
    int rank = 0;
    for (int i = 0; i < pivots; ++i)
        if (fabs(P_(i, i)) > cut)
            ++rank;
    double h[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (rank > 0)
    {
        // rhs = e_{R-1}; only the row swaps move its single 1, and only y[0..K) is used afterwards (the rows below
        // the square part never reach the solution)
        int pos = R - 1;
        for (int k = 0; k < K; ++k)
        {
            if (pos == k)
                pos = row_swap[k];
            else if (pos == row_swap[k])
                pos = k;
        }
        double y[9];
        for (int i = 0; i < 9; i++)
            y[i] = i == pos ? 1.0 : 0.0;
        for (int c = 0; c < K; ++c)
        {
            const double yc = y[c];
            if (yc == 0.0)
                continue;
            for (int r = c + 1; r < K; ++r)
                y[r] = __dsub_rn(y[r], __dmul_rn(yc, P_(r, c)));
        }
        for (int c = rank - 1; c >= 0; --c)
        {
            if (y[c] == 0.0)
                continue;
            y[c] = __ddiv_rn(y[c], P_(c, c));
            const double yc = y[c];
            for (int r = 0; r < c; ++r)
                y[r] = __dsub_rn(y[r], __dmul_rn(yc, P_(r, c)));
        }
        for (int i = 0; i < rank; ++i)
            h[i] = y[i];
        for (int k = K - 1; k >= 0; --k)
        {
            const double tmp = h[k];
            h[k] = h[col_swap[k]];
            h[col_swap[k]] = tmp;
        }
    }
    double M18[18];
    finish_model(h, M18);
    for (int i = 0; i < 18; i++)
        job.model_out[i] = M18[i];
#undef P_
}

size_t k3_inlier_scratch_bytes(size_t n)
{
    return (2 * n + 1) * 9 * sizeof(double);
}

int k3_fit_inliers(const K3InlierJob *d_jobs, size_t n_jobs, cudaStream_t stream)
{
    if (n_jobs == 0)
        return 0;
    k3_fit_inliers_kernel<<<(unsigned)n_jobs, K3I_THREADS, 0, stream>>>(d_jobs);
    count_launch();
    OCB_CUDA(cudaGetLastError());
    return 0;
}

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
