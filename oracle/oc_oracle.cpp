// TEST INFRASTRUCTURE ONLY -- see oc_oracle.hpp. CPU restatement of the reference hot path.
// Build with -ffp-contract=off (the reference ships x86-64 baseline flags: no FMA, CMakeLists.txt:59).
#include "oc_oracle.hpp"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <limits>
#include <numeric>
#include <random>
#include <unordered_map>

namespace oc_oracle
{

int minimum_points(int kind)
{
    return kind == MODEL_HOMOGRAPHY ? 4 : (kind == MODEL_ESSENTIAL ? 5 : 8);
}
double default_threshold(int kind)
{
    return kind == MODEL_HOMOGRAPHY ? 0.005 : 0.01;
}
Model::Model()
{
    for (int i = 0; i < 9; i++)
        M[i] = Minv[i] = NAN; // *_model.cpp ctors: Matrix3d::Constant(NAN)
}
Model::Model(int kind_) : Model()
{
    kind = kind_;
    thr = default_threshold(kind_);
}

// =================================================================================================
// src/match
// =================================================================================================

static inline int hamming512(const uint64_t *a, const uint64_t *b)
{
    // std::bitset<486>::operator^ + count() = sum of popcountl over the 8 words
    // (/usr/include/c++/13/bitset:230-234); bits 486..511 are zero in both operands.
    int d = 0;
    for (int w = 0; w < DESCRIPTOR_WORDS; w++)
        d += __builtin_popcountll(a[w] ^ b[w]);
    return d;
}

void match_top2(const uint64_t *q, size_t n1, const uint64_t *c, size_t n2, uint32_t *best_k, uint16_t *best_d,
                uint16_t *second_d)
{
    // src/match/match_features.cpp:71-93 with integer distances: d*(1.0/486) is strictly monotone
    // in d on [0,486], so the double comparisons of the reference order exactly like these.
    // Queries are independent (the reference's outer loop :71 carries no state from one query to the next), so the
    // checker runs them on all host cores: full-size parity checks then take seconds.
    const int INF = 0xFFFF;
#pragma omp parallel for schedule(static) if (n1 * n2 > (size_t)1 << 22)
    for (size_t i = 0; i < n1; i++)
    {
        int best = INF, second = INF;
        uint32_t bk = 0; // feature_match best_match{i, 0, inf}: index 0 when nothing was seen (:74)
        const uint64_t *qi = q + i * DESCRIPTOR_WORDS;
        for (size_t k = 0; k < n2; k++)
        {
            int d = hamming512(qi, c + k * DESCRIPTOR_WORDS);
            if (d < second) // :80
            {
                if (d < best) // :82
                {
                    second = best;
                    best = d;
                    bk = (uint32_t)k;
                }
                else
                {
                    second = d; // :90 -- a later equal distance becomes second best
                }
            }
        }
        best_k[i] = bk;
        best_d[i] = (uint16_t)best;
        second_d[i] = (uint16_t)second;
    }
}

void match_lists_top2(const uint64_t *q, const uint64_t *c, const uint32_t *list_query, const uint64_t *list_begin,
                      const uint32_t *list_candidates, size_t n_lists, uint32_t *best_pos, double *best_dist,
                      double *second_dist)
{
    for (size_t l = 0; l < n_lists; l++)
    {
        // src/dense/dense_stereo.cpp:251-273
        double best = std::numeric_limits<double>::infinity();
        double second = std::numeric_limits<double>::infinity();
        uint32_t bp = 0;
        const uint64_t *ql = q + (size_t)list_query[l] * DESCRIPTOR_WORDS;
        for (uint64_t k = list_begin[l]; k < list_begin[l + 1]; k++)
        {
            // descriptor_distance, :56-59
            const double d = hamming512(ql, c + (size_t)list_candidates[k] * DESCRIPTOR_WORDS) * (1.0 / DESCRIPTOR_BITS);
            if (d < second) // :258
            {
                if (d < best) // :260
                {
                    second = best;
                    best = d;
                    bp = (uint32_t)(k - list_begin[l]);
                }
                else
                {
                    second = d; // :268
                }
            }
        }
        best_pos[l] = bp;
        best_dist[l] = best;
        second_dist[l] = second;
    }
}

bool guided_good_match(size_t list_length, double best_dist, double second_dist)
{
    constexpr double RATIO_THRESHOLD = 0.85;                  // src/dense/dense_stereo.cpp:51
    constexpr double MAX_ABSOLUTE_DESCRIPTOR_DISTANCE = 0.35; // :53
    if (list_length == 0)
        return false; // :248-249
    return list_length >= 2 ? best_dist < RATIO_THRESHOLD * second_dist : best_dist < MAX_ABSOLUTE_DESCRIPTOR_DISTANCE;
}

void match_col_best(const uint64_t *q, size_t n1, const uint64_t *c, size_t n2, uint32_t *col_best_q)
{
#pragma omp parallel for schedule(static) if (n1 * n2 > (size_t)1 << 22)
    for (size_t k = 0; k < n2; k++)
    {
        int best = 0x7FFFFFFF;
        uint32_t bq = 0xFFFFFFFFu;
        for (size_t i = 0; i < n1; i++)
        {
            int d = hamming512(q + i * DESCRIPTOR_WORDS, c + k * DESCRIPTOR_WORDS);
            if (d < best)
            {
                best = d;
                bq = (uint32_t)i;
            }
        }
        col_best_q[k] = bq;
    }
}

std::vector<Match> match_features_subset(const uint64_t *desc1, const uint64_t *desc2, const size_t *idx1, size_t n1,
                                         const size_t *idx2, size_t n2)
{
    // src/match/match_features.cpp:62-66 -- pack set_2's subset contiguously
    std::vector<uint64_t> packed_2(n2 * DESCRIPTOR_WORDS);
    for (size_t k = 0; k < n2; k++)
        std::memcpy(&packed_2[k * DESCRIPTOR_WORDS], desc2 + idx2[k] * DESCRIPTOR_WORDS, 64);

    std::vector<Match> results;
    results.reserve(n1);
    const double inf = std::numeric_limits<double>::infinity();
    for (size_t a = 0; a < n1; a++) // :71 for (size_t i : indices_1)
    {
        const size_t i = idx1[a];
        const uint64_t *d1 = desc1 + i * DESCRIPTOR_WORDS;
        Match best{i, 0, inf}; // :74
        double second_best = inf;
        for (size_t k = 0; k < n2; k++)
        {
            double distance = hamming512(d1, &packed_2[k * DESCRIPTOR_WORDS]) * (1.0 / DESCRIPTOR_BITS); // :79
            if (distance < second_best)
            {
                if (distance < best.distance)
                {
                    second_best = best.distance;
                    best.distance = distance;
                    best.feature_index_2 = idx2[k];
                }
                else
                {
                    second_best = distance;
                }
            }
        }
        if (best.distance < 0.8 * second_best) // :94
            results.push_back(best);
    }
    // :100-101 -- descending, unstable; same std::sort, same comparator results => same order
    std::sort(results.begin(), results.end(), [](const Match &f1, const Match &f2) { return f1.distance > f2.distance; });
    return results;
}

std::vector<size_t> spatially_subsample_feature_indices(const double *xy, const float *strength, size_t n_features,
                                                        double spacing_pixels, size_t count)
{
    // src/match/match_features.cpp:8-52. The reference asks a jk KD-tree for the nearest kept point
    // (SquaredL2 = dx*dx + dy*dy, external/jk-tree/include/jk/KDTree.h:681-690) and keeps the feature iff
    // that squared distance > spacing^2; the minimum over kept points is computed here by brute force
    // over a uniform grid (cell = spacing), which yields the same minimum-or-above-threshold decision.
    if (count == 0)
        count = n_features;
    if (count == 0)
        return {};
    std::vector<size_t> sorted_indices(count);
    for (size_t i = 0; i < count; i++)
        sorted_indices[i] = i;
    std::sort(sorted_indices.begin(), sorted_indices.end(),
              [strength](size_t a, size_t b) { return strength[a] > strength[b]; }); // :22-23

    std::vector<size_t> indices;
    indices.reserve(n_features / 4);
    const double thr = spacing_pixels * spacing_pixels;
    const bool use_grid = spacing_pixels > 0 && std::isfinite(spacing_pixels);
    std::unordered_map<uint64_t, std::vector<size_t>> grid;
    auto cell_of = [&](double v) -> int64_t { return (int64_t)std::floor(v / spacing_pixels); };
    auto key_of = [](int64_t cx, int64_t cy) -> uint64_t {
        return ((uint64_t)(uint32_t)(int32_t)cx << 32) | (uint64_t)(uint32_t)(int32_t)cy;
    };
    for (size_t idx : sorted_indices)
    {
        const double x = xy[2 * idx], y = xy[2 * idx + 1];
        bool keep = true;
        if (!indices.empty())
        {
            if (use_grid && std::isfinite(x) && std::isfinite(y))
            {
                const int64_t cx = cell_of(x), cy = cell_of(y);
                for (int64_t gx = cx - 2; gx <= cx + 2 && keep; gx++) // +-2 cells: immune to floor() rounding
                    for (int64_t gy = cy - 2; gy <= cy + 2 && keep; gy++)
                    {
                        auto it = grid.find(key_of(gx, gy));
                        if (it == grid.end())
                            continue;
                        for (size_t j : it->second)
                        {
                            const double dx = x - xy[2 * j], dy = y - xy[2 * j + 1];
                            if (!(dx * dx + dy * dy > thr)) // :45 keep iff nn distance > spacing^2
                            {
                                keep = false;
                                break;
                            }
                        }
                    }
            }
            else
            {
                for (size_t j : indices)
                {
                    const double dx = x - xy[2 * j], dy = y - xy[2 * j + 1];
                    if (!(dx * dx + dy * dy > thr))
                    {
                        keep = false;
                        break;
                    }
                }
            }
        }
        if (keep)
        {
            if (use_grid && std::isfinite(x) && std::isfinite(y))
                grid[key_of(cell_of(x), cell_of(y))].push_back(idx);
            indices.push_back(idx);
        }
    }
    return indices;
}

// =================================================================================================
// Linear algebra restated from Eigen 3.4.0 (column-major storage: A(r,c) = A[r + rows*c])
// =================================================================================================
namespace la
{

void fullpivlu_solve(const double *A, int rows, int cols, const double *b, double *x)
{
    // Eigen/src/LU/FullPivLU.h computeInPlace + _solve_impl.
    const int size = std::min(rows, cols);
    std::vector<double> lu(A, A + (size_t)rows * cols);
    auto LU = [&](int r, int c) -> double & { return lu[(size_t)r + (size_t)rows * c]; };
    std::vector<int> rowT(size), colT(size);
    int nonzero_pivots = size;
    double maxpivot = 0;
    for (int k = 0; k < size; k++)
    {
        // maxCoeff visitor: column-major traversal of the bottom-right corner, strict '>' (first max wins)
        int br = k, bc = k;
        double biggest = std::fabs(LU(k, k));
        for (int j = k; j < cols; j++)
            for (int i = k; i < rows; i++)
            {
                const double v = std::fabs(LU(i, j));
                if (v > biggest)
                {
                    biggest = v;
                    br = i;
                    bc = j;
                }
            }
        if (biggest == 0)
        {
            nonzero_pivots = k;
            for (int i = k; i < size; i++)
            {
                rowT[i] = i;
                colT[i] = i;
            }
            break;
        }
        if (biggest > maxpivot)
            maxpivot = biggest;
        rowT[k] = br;
        colT[k] = bc;
        if (k != br)
            for (int j = 0; j < cols; j++)
                std::swap(LU(k, j), LU(br, j));
        if (k != bc)
            for (int i = 0; i < rows; i++)
                std::swap(LU(i, k), LU(i, bc));
        if (k < rows - 1)
        {
            const double piv = LU(k, k);
            for (int i = k + 1; i < rows; i++)
                LU(i, k) /= piv;
        }
        if (k < size - 1)
            for (int j = k + 1; j < cols; j++)
            {
                const double u = LU(k, j);
                for (int i = k + 1; i < rows; i++)
                    LU(i, j) -= LU(i, k) * u;
            }
    }
    // rank(): pivots with |lu(i,i)| > threshold*|maxpivot|, threshold = eps * diagonalSize
    const double premult = std::fabs(maxpivot) * (DBL_EPSILON * (double)size);
    int rank = 0;
    for (int i = 0; i < nonzero_pivots; i++)
        rank += (std::fabs(LU(i, i)) > premult) ? 1 : 0;
    if (rank == 0)
    {
        for (int i = 0; i < cols; i++)
            x[i] = 0;
        return;
    }
    // Step 1: c = P * rhs
    std::vector<double> c(b, b + rows);
    for (int k = 0; k < size; k++)
        std::swap(c[k], c[rowT[k]]);
    // Step 2: unit-lower solve on the top smalldim rows (column-oriented), then the rows below
    for (int i = 0; i < size; i++)
    {
        const double ci = c[i];
        if (ci != 0)
            for (int r = i + 1; r < size; r++)
                c[r] -= ci * LU(r, i);
    }
    if (rows > cols)
        for (int r = cols; r < rows; r++)
        {
            double acc = 0;
            for (int j = 0; j < cols; j++)
                acc += LU(r, j) * c[j];
            c[r] -= acc;
        }
    // Step 3: upper solve on the leading rank x rank block (column-oriented back substitution)
    for (int i = rank - 1; i >= 0; i--)
    {
        if (c[i] != 0)
        {
            c[i] /= LU(i, i);
            const double ci = c[i];
            for (int r = 0; r < i; r++)
                c[r] -= ci * LU(r, i);
        }
    }
    // Step 4: x = Q * [c(0..rank-1); 0]
    std::vector<double> y(cols, 0.0);
    for (int i = 0; i < rank; i++)
        y[i] = c[i];
    for (int k = size - 1; k >= 0; k--)
        std::swap(y[k], y[colT[k]]);
    for (int i = 0; i < cols; i++)
        x[i] = y[i];
}

void inverse3(const double *M, double *Minv)
{
    // Eigen/src/LU/InverseImpl.h compute_inverse_size3_helper: cofactors, det from column 0.
    auto m = [&](int r, int c) { return M[r + 3 * c]; };
    auto cof = [&](int i, int j) {
        const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
        return m(i1, j1) * m(i2, j2) - m(i1, j2) * m(i2, j1);
    };
    const double c00 = cof(0, 0), c10 = cof(1, 0), c20 = cof(2, 0);
    const double det = (c00 * m(0, 0) + c10 * m(1, 0)) + c20 * m(2, 0);
    const double invdet = 1.0 / det;
    auto out = [&](int r, int c) -> double & { return Minv[r + 3 * c]; };
    out(0, 0) = c00 * invdet;
    out(0, 1) = c10 * invdet;
    out(0, 2) = c20 * invdet;
    out(1, 0) = cof(0, 1) * invdet;
    out(1, 1) = cof(1, 1) * invdet;
    out(1, 2) = cof(2, 1) * invdet;
    out(2, 0) = cof(0, 2) * invdet;
    out(2, 1) = cof(1, 2) * invdet;
    out(2, 2) = cof(2, 2) * invdet;
}

namespace
{
struct Rot // Eigen::JacobiRotation<double>
{
    double c, s;
};
inline Rot rot_transpose(Rot r)
{
    return Rot{r.c, -r.s};
}
inline Rot rot_mul(Rot a, Rot b) // JacobiRotation::operator*
{
    return Rot{a.c * b.c - a.s * b.s, a.c * b.s + a.s * b.c};
}
// apply_rotation_in_the_plane on two strided vectors
inline void apply_rot(double *x, int incx, double *y, int incy, int n, Rot j)
{
    if (j.c == 1 && j.s == 0)
        return;
    for (int i = 0; i < n; i++)
    {
        const double xi = x[i * incx], yi = y[i * incy];
        x[i * incx] = j.c * xi + j.s * yi;
        y[i * incy] = -j.s * xi + j.c * yi;
    }
}
inline bool make_jacobi(double x, double y, double z, Rot &r) // JacobiRotation::makeJacobi(x,y,z)
{
    const double deno = 2.0 * std::fabs(y);
    if (deno < DBL_MIN)
    {
        r.c = 1;
        r.s = 0;
        return false;
    }
    const double tau = (x - z) / deno;
    const double w = std::sqrt(tau * tau + 1.0);
    const double t = tau > 0 ? 1.0 / (tau + w) : 1.0 / (tau - w);
    const double sign_t = t > 0 ? 1.0 : -1.0;
    const double n = 1.0 / std::sqrt(t * t + 1.0);
    r.s = -sign_t * (y / std::fabs(y)) * std::fabs(t) * n;
    r.c = n;
    return true;
}
// Eigen/src/SVD/JacobiSVD.h real_2x2_jacobi_svd on W(p,p),W(p,q),W(q,p),W(q,q)
inline void real_2x2_jacobi_svd(double mpp, double mpq, double mqp, double mqq, Rot &j_left, Rot &j_right)
{
    double m00 = mpp, m01 = mpq, m10 = mqp, m11 = mqq;
    Rot rot1;
    const double t = m00 + m11;
    const double d = m10 - m01;
    if (std::fabs(d) < DBL_MIN)
    {
        rot1.s = 0;
        rot1.c = 1;
    }
    else
    {
        const double u = t / d;
        const double tmp = std::sqrt(1.0 + u * u);
        rot1.s = 1.0 / tmp;
        rot1.c = u / tmp;
    }
    // m.applyOnTheLeft(0,1,rot1)
    {
        const double a0 = m00, a1 = m01, b0 = m10, b1 = m11;
        m00 = rot1.c * a0 + rot1.s * b0;
        m01 = rot1.c * a1 + rot1.s * b1;
        m10 = -rot1.s * a0 + rot1.c * b0;
        m11 = -rot1.s * a1 + rot1.c * b1;
    }
    (void)m10;
    make_jacobi(m00, m01, m11, j_right);
    j_left = rot_mul(rot1, rot_transpose(j_right));
}
} // namespace

void jacobi_svd_square(const double *A, int n, double *U, double *S, double *V)
{
    // Eigen/src/SVD/JacobiSVD.h compute() for a square real matrix (no QR preconditioner).
    std::vector<double> W((size_t)n * n), Um, Vm;
    double scale = 0;
    for (int i = 0; i < n * n; i++)
    {
        const double a = std::fabs(A[i]);
        if (a > scale)
            scale = a;
    }
    if (!(scale > 0) || !std::isfinite(scale))
        scale = 1.0; // Eigen: if(!(numext::isfinite)(scale)) -> invalid input; scale==0 -> 1
    for (int i = 0; i < n * n; i++)
        W[i] = A[i] / scale;
    auto w = [&](int r, int c) -> double & { return W[(size_t)r + (size_t)n * c]; };
    if (U)
    {
        Um.assign((size_t)n * n, 0.0);
        for (int i = 0; i < n; i++)
            Um[(size_t)i + (size_t)n * i] = 1.0;
    }
    if (V)
    {
        Vm.assign((size_t)n * n, 0.0);
        for (int i = 0; i < n; i++)
            Vm[(size_t)i + (size_t)n * i] = 1.0;
    }
    const double precision = 2.0 * DBL_EPSILON;
    const double considerAsZero = DBL_MIN;
    double maxDiag = 0;
    for (int i = 0; i < n; i++)
        maxDiag = std::max(maxDiag, std::fabs(w(i, i)));
    bool finished = false;
    int guard = 0;
    while (!finished && guard++ < 1000)
    {
        finished = true;
        for (int p = 1; p < n; p++)
            for (int q = 0; q < p; q++)
            {
                const double threshold = std::max(considerAsZero, precision * maxDiag);
                if (std::fabs(w(p, q)) > threshold || std::fabs(w(q, p)) > threshold)
                {
                    finished = false;
                    Rot jl, jr;
                    real_2x2_jacobi_svd(w(p, p), w(p, q), w(q, p), w(q, q), jl, jr);
                    // W.applyOnTheLeft(p,q,j_left): rows p and q
                    apply_rot(&w(p, 0), n, &w(q, 0), n, n, jl);
                    if (U) // U.applyOnTheRight(p,q,j_left.transpose()): columns p,q with j^T^T = j
                        apply_rot(&Um[(size_t)n * p], 1, &Um[(size_t)n * q], 1, n, jl);
                    // W.applyOnTheRight(p,q,j_right): columns p,q with j.transpose()
                    apply_rot(&w(0, p), 1, &w(0, q), 1, n, rot_transpose(jr));
                    if (V)
                        apply_rot(&Vm[(size_t)n * p], 1, &Vm[(size_t)n * q], 1, n, rot_transpose(jr));
                    maxDiag = std::max(maxDiag, std::max(std::fabs(w(p, p)), std::fabs(w(q, q))));
                }
            }
    }
    // positive singular values
    for (int i = 0; i < n; i++)
    {
        const double a = std::fabs(w(i, i));
        S[i] = a;
        if (U && a != 0)
        {
            const double sgn = w(i, i) / a;
            for (int r = 0; r < n; r++)
                Um[(size_t)r + (size_t)n * i] *= sgn;
        }
    }
    for (int i = 0; i < n; i++)
        S[i] *= scale;
    // sort descending (selection, swapping columns)
    for (int i = 0; i < n; i++)
    {
        int pos = 0;
        double mx = S[i];
        for (int k = 1; k < n - i; k++)
            if (S[i + k] > mx)
            {
                mx = S[i + k];
                pos = k;
            }
        if (mx == 0)
            break;
        if (pos)
        {
            pos += i;
            std::swap(S[i], S[pos]);
            if (U)
                for (int r = 0; r < n; r++)
                    std::swap(Um[(size_t)r + (size_t)n * i], Um[(size_t)r + (size_t)n * pos]);
            if (V)
                for (int r = 0; r < n; r++)
                    std::swap(Vm[(size_t)r + (size_t)n * i], Vm[(size_t)r + (size_t)n * pos]);
        }
    }
    if (U)
        std::memcpy(U, Um.data(), sizeof(double) * n * n);
    if (V)
        std::memcpy(V, Vm.data(), sizeof(double) * n * n);
}

void jacobi_svd_tall_v(const double *A, int rows, int cols, double *S, double *V)
{
    // Eigen preconditions a rows>cols JacobiSVD with a column-pivoting Householder QR and runs the
    // Jacobi sweeps on R. Only V and S are needed by fundamental_matrix_model.cpp:188-189, so this
    // restatement does the same: A*P = Q*R (Businger-Golub pivoting), R = U' S W^T, V = P*W.
    std::vector<double> a(A, A + (size_t)rows * cols);
    auto at = [&](int r, int c) -> double & { return a[(size_t)r + (size_t)rows * c]; };
    std::vector<int> perm(cols);
    std::iota(perm.begin(), perm.end(), 0);
    const int steps = std::min(rows, cols);
    for (int k = 0; k < steps; k++)
    {
        int piv = k;
        double best = -1;
        for (int j = k; j < cols; j++)
        {
            double s = 0;
            for (int i = k; i < rows; i++)
                s += at(i, j) * at(i, j);
            if (s > best)
            {
                best = s;
                piv = j;
            }
        }
        if (piv != k)
        {
            for (int i = 0; i < rows; i++)
                std::swap(at(i, k), at(i, piv));
            std::swap(perm[k], perm[piv]);
        }
        // Householder on column k, rows k..rows-1
        double tail = 0;
        for (int i = k + 1; i < rows; i++)
            tail += at(i, k) * at(i, k);
        const double c0 = at(k, k);
        if (tail <= DBL_MIN)
            continue;
        double beta = std::sqrt(c0 * c0 + tail);
        if (c0 >= 0)
            beta = -beta;
        const double tau = (beta - c0) / beta;
        std::vector<double> v(rows - k);
        v[0] = 1.0;
        for (int i = k + 1; i < rows; i++)
            v[i - k] = at(i, k) / (c0 - beta);
        for (int j = k; j < cols; j++)
        {
            double dot = 0;
            for (int i = k; i < rows; i++)
                dot += v[i - k] * at(i, j);
            dot *= tau;
            for (int i = k; i < rows; i++)
                at(i, j) -= dot * v[i - k];
        }
    }
    std::vector<double> R((size_t)cols * cols, 0.0), W((size_t)cols * cols);
    for (int j = 0; j < cols; j++)
        for (int i = 0; i <= j && i < rows; i++)
            R[(size_t)i + (size_t)cols * j] = at(i, j);
    jacobi_svd_square(R.data(), cols, nullptr, S, W.data());
    for (int j = 0; j < cols; j++)
        for (int i = 0; i < cols; i++)
            V[(size_t)perm[i] + (size_t)cols * j] = W[(size_t)i + (size_t)cols * j];
}

} // namespace la

// =================================================================================================
// src/model_inliers: residuals (canonical operation order)
// =================================================================================================

static inline double h_error(const double *H, const double *G, double x1, double y1, double x2, double y2)
{
    // homography_model.cpp:89-97. m = meas / meas.z => m.z == 1.0 exactly, so H(r,2)*m.z == H(r,2).
    // Dot products left to right: (H(r,0)*x + H(r,1)*y) + H(r,2).
    const double px = (H[0] * x1 + H[3] * y1) + H[6];
    const double py = (H[1] * x1 + H[4] * y1) + H[7];
    const double pz = (H[2] * x1 + H[5] * y1) + H[8];
    const double dx = px / pz - x2, dy = py / pz - y2;
    const double fwd = dx * dx + dy * dy;
    const double qx = (G[0] * x2 + G[3] * y2) + G[6];
    const double qy = (G[1] * x2 + G[4] * y2) + G[7];
    const double qz = (G[2] * x2 + G[5] * y2) + G[8];
    const double ex = qx / qz - x1, ey = qy / qz - y1;
    const double bwd = ex * ex + ey * ey;
    return std::sqrt((fwd + bwd) / 2.0);
}

static inline double epi_error(const double *E, double x1, double y1, double x2, double y2)
{
    // essential_matrix_model.cpp:112-123 == fundamental_matrix_model.cpp:110-121.
    // (x2^T E) is formed first, then dotted with x1; E^T x2 is the same row vector.
    const double b0 = (x2 * E[0] + y2 * E[1]) + E[2];
    const double b1 = (x2 * E[3] + y2 * E[4]) + E[5];
    const double b2 = (x2 * E[6] + y2 * E[7]) + E[8];
    const double r = (b0 * x1 + b1 * y1) + b2;
    const double a0 = (E[0] * x1 + E[3] * y1) + E[6];
    const double a1 = (E[1] * x1 + E[4] * y1) + E[7];
    const double denom = ((a0 * a0 + a1 * a1) + b0 * b0) + b1 * b1;
    if (denom < 1e-20)
        return std::numeric_limits<double>::max();
    return std::sqrt((r * r) / denom);
}

// m / m.z for both measurements; NaN when z/z != 1 (z zero, infinite or NaN), which is what the
// reference's m.z component then is and what poisons every term of its residual.
static inline void normalise(const Corr &c, double &x1, double &y1, double &x2, double &y2)
{
    x1 = c.m1[0] / c.m1[2];
    y1 = c.m1[1] / c.m1[2];
    x2 = c.m2[0] / c.m2[2];
    y2 = c.m2[1] / c.m2[2];
    if (!(c.m1[2] / c.m1[2] == 1.0))
        x1 = y1 = NAN;
    if (!(c.m2[2] / c.m2[2] == 1.0))
        x2 = y2 = NAN;
}

double error(const Model &m, const Corr &c)
{
    double x1, y1, x2, y2;
    normalise(c, x1, y1, x2, y2);
    if (m.kind == MODEL_HOMOGRAPHY)
        return h_error(m.M, m.Minv, x1, y1, x2, y2);
    return epi_error(m.M, x1, y1, x2, y2);
}

double evaluate(const Model &m, const Corr *c, size_t n, std::vector<bool> &inliers)
{
    // homography_model.cpp:99-118 and twins
    inliers.resize(n);
    double total_score = 0;
    for (size_t i = 0; i < n; i++)
    {
        const double e = error(m, c[i]);
        if (e < m.thr)
        {
            inliers[i] = true;
            const double ratio = e / m.thr;
            total_score += 1.0 - ratio * ratio;
        }
        else
        {
            inliers[i] = false;
        }
    }
    return total_score;
}

void score_hypothesis(const Model &m, const Corr *c, size_t n, const size_t *order, double *score, uint32_t *count,
                      uint32_t *bits)
{
    double s = 0;
    uint32_t cnt = 0;
    if (bits)
        std::memset(bits, 0, sizeof(uint32_t) * ((n + 31) / 32));
    for (size_t p = 0; p < n; p++)
    {
        const size_t idx = order ? order[p] : p;
        const double e = error(m, c[idx]);
        if (e < m.thr) // ransac.cpp:189-194
        {
            if (bits)
                bits[idx >> 5] |= 1u << (idx & 31);
            cnt++;
            const double ratio = e / m.thr;
            s += 1.0 - ratio * ratio;
        }
    }
    *score = s;
    *count = cnt;
}

// =================================================================================================
// src/model_inliers: fits
// =================================================================================================

static void set_h_from_solution(Model &m, const double *h)
{
    // homography.row(r) = H_.segment(3r,3); homography /= homography(2,2); inverse()
    // (homography_model.cpp:45-49)
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++)
            m.M[r + 3 * c] = h[3 * r + c];
    const double h22 = m.M[8];
    for (int i = 0; i < 9; i++)
        m.M[i] /= h22;
    la::inverse3(m.M, m.Minv);
}

static void h_rows(const Corr &c, double *r0, double *r1)
{
    // homography_model.cpp:26-35
    const double x = c.m1[0] / c.m1[2], y = c.m1[1] / c.m1[2];
    const double x_ = c.m2[0] / c.m2[2], y_ = c.m2[1] / c.m2[2];
    const double a[9] = {-x, -y, -1, 0, 0, 0, x * x_, y * x_, x_};
    const double b[9] = {0, 0, 0, -x, -y, -1, x * y_, y * y_, y_};
    std::memcpy(r0, a, sizeof a);
    std::memcpy(r1, b, sizeof b);
}

static void epi_row(const Corr &c, double *r)
{
    // essential_matrix_model.cpp:52-59
    const double x = c.m1[0] / c.m1[2], y = c.m1[1] / c.m1[2];
    const double x_ = c.m2[0] / c.m2[2], y_ = c.m2[1] / c.m2[2];
    const double a[9] = {x * x_, x * y_, x, y * x_, y * y_, y, x_, y_, 1};
    std::memcpy(r, a, sizeof a);
}

static void fit_h_rows(Model &m, const std::vector<std::array<double, 9>> &rows_in)
{
    // rows_in: the 2k DLT rows; append the h33 = 1 constraint row (homography_model.cpp:37-44 / :76-81)
    const int rows = (int)rows_in.size() + 1;
    std::vector<double> P((size_t)rows * 9, 0.0), rhs(rows, 0.0);
    for (int r = 0; r < rows - 1; r++)
        for (int c = 0; c < 9; c++)
            P[(size_t)r + (size_t)rows * c] = rows_in[r][c];
    P[(size_t)(rows - 1) + (size_t)rows * 8] = 1.0;
    rhs[rows - 1] = 1.0;
    double h[9];
    la::fullpivlu_solve(P.data(), rows, 9, rhs.data(), h);
    set_h_from_solution(m, h);
}

static void fit_epipolar_rows(Model &m, const std::vector<std::array<double, 9>> &A)
{
    // calculateEssentialMatrix / calculateFundamentalMatrix
    // (essential_matrix_model.cpp:12-31, fundamental_matrix_model.cpp:13-29)
    double AtA[81];
    for (int j = 0; j < 9; j++)
        for (int i = 0; i < 9; i++)
        {
            double s = 0;
            for (size_t r = 0; r < A.size(); r++)
                s += A[r][i] * A[r][j];
            AtA[i + 9 * j] = s;
        }
    double S9[9], V9[81];
    la::jacobi_svd_square(AtA, 9, nullptr, S9, V9);
    double F[9]; // column-major 3x3; F.row(r) = last column of V, entries 3r..3r+2
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++)
            F[r + 3 * c] = V9[(3 * r + c) + 9 * 8];
    double U[9], S[3], V[9];
    la::jacobi_svd_square(F, 3, U, S, V);
    if (m.kind == MODEL_ESSENTIAL)
    {
        const double avg = (S[0] + S[1]) / 2.0;
        S[0] = avg;
        S[1] = avg;
    }
    S[2] = 0;
    // (U * diag(S)) * V^T
    for (int c = 0; c < 3; c++)
        for (int r = 0; r < 3; r++)
        {
            const double t0 = (U[r + 0] * S[0]) * V[c + 0];
            const double t1 = (U[r + 3] * S[1]) * V[c + 3];
            const double t2 = (U[r + 6] * S[2]) * V[c + 6];
            m.M[r + 3 * c] = (t0 + t1) + t2;
        }
}

void fit(Model &m, const Corr *c, const size_t *sample)
{
    const int k = minimum_points(m.kind);
    std::vector<std::array<double, 9>> rows;
    if (m.kind == MODEL_HOMOGRAPHY)
    {
        rows.resize(8);
        for (int i = 0; i < 4; i++)
            h_rows(c[sample[i]], rows[2 * i].data(), rows[2 * i + 1].data());
        fit_h_rows(m, rows);
    }
    else
    {
        rows.resize(k);
        for (int i = 0; i < k; i++)
            epi_row(c[sample[i]], rows[i].data());
        fit_epipolar_rows(m, rows);
    }
}

void fit_inliers(Model &m, const Corr *c, size_t n, const std::vector<bool> &inliers)
{
    const size_t num = (size_t)std::count(inliers.begin(), inliers.end(), true);
    std::vector<std::array<double, 9>> rows;
    if (m.kind == MODEL_HOMOGRAPHY)
    {
        // homography_model.cpp:52-87 (no minimum-count guard in the reference)
        rows.resize(num * 2);
        for (size_t i = 0, j = 0; i < n; i++)
            if (inliers[i])
            {
                h_rows(c[i], rows[2 * j].data(), rows[2 * j + 1].data());
                j++;
            }
        fit_h_rows(m, rows);
    }
    else
    {
        if (num < (size_t)minimum_points(m.kind)) // essential_matrix_model.cpp:65-66
            return;
        rows.resize(num);
        for (size_t i = 0, j = 0; i < n; i++)
            if (inliers[i])
                epi_row(c[i], rows[j++].data());
        fit_epipolar_rows(m, rows);
    }
}

bool check_sample_degeneracy_h(const Corr *c, const size_t *sample)
{
    // homography_model.cpp:120-136
    double px[4], py[4];
    for (int i = 0; i < 4; i++)
    {
        px[i] = c[sample[i]].m1[0] / c[sample[i]].m1[2];
        py[i] = c[sample[i]].m1[1] / c[sample[i]].m1[2];
    }
    for (int i = 0; i < 4; i++)
        for (int j = i + 1; j < 4; j++)
            for (int k = j + 1; k < 4; k++)
            {
                const double v1x = px[j] - px[i], v1y = py[j] - py[i];
                const double v2x = px[k] - px[i], v2y = py[k] - py[i];
                if (std::abs(v1x * v2y - v1y * v2x) < 1e-10)
                    return true;
            }
    return false;
}

void check_degeneracy_f(Model &m, const Corr *c, size_t n, std::vector<bool> &inliers)
{
    // fundamental_matrix_model.cpp:123-215 (DEGENSAC)
    std::vector<size_t> f_idx;
    for (size_t i = 0; i < inliers.size(); i++)
        if (inliers[i])
            f_idx.push_back(i);
    if (f_idx.size() < 4)
        return;
    Model h(MODEL_HOMOGRAPHY);
    h.thr = m.thr * 2;
    size_t h_indices[4];
    for (int i = 0; i < 4; i++)
        h_indices[i] = f_idx[i];
    fit(h, c, h_indices);

    std::vector<bool> h_inl(n, false);
    size_t h_count = 0;
    for (size_t idx : f_idx)
        if (error(h, c[idx]) < h.thr)
        {
            h_inl[idx] = true;
            h_count++;
        }
    const double h_ratio = static_cast<double>(h_count) / f_idx.size();
    if (h_ratio < 0.7)
        return;
    fit_inliers(h, c, n, h_inl);
    std::vector<size_t> non_h;
    for (size_t idx : f_idx)
    {
        if (error(h, c[idx]) < h.thr)
            h_inl[idx] = true;
        else
        {
            h_inl[idx] = false;
            non_h.push_back(idx);
        }
    }
    if (non_h.size() < 2)
        return;
    // A.row(i) = x2.cross(H * x1)
    const int rows = (int)non_h.size();
    std::vector<double> A((size_t)rows * 3);
    for (int i = 0; i < rows; i++)
    {
        const Corr &cc = c[non_h[i]];
        const double x1[3] = {cc.m1[0] / cc.m1[2], cc.m1[1] / cc.m1[2], cc.m1[2] / cc.m1[2]};
        const double x2[3] = {cc.m2[0] / cc.m2[2], cc.m2[1] / cc.m2[2], cc.m2[2] / cc.m2[2]};
        double hx[3];
        for (int r = 0; r < 3; r++)
            hx[r] = (h.M[r] * x1[0] + h.M[r + 3] * x1[1]) + h.M[r + 6] * x1[2];
        A[(size_t)i + (size_t)rows * 0] = x2[1] * hx[2] - x2[2] * hx[1];
        A[(size_t)i + (size_t)rows * 1] = x2[2] * hx[0] - x2[0] * hx[2];
        A[(size_t)i + (size_t)rows * 2] = x2[0] * hx[1] - x2[1] * hx[0];
    }
    double S3[3], V3[9];
    if (rows >= 3)
        la::jacobi_svd_tall_v(A.data(), rows, 3, S3, V3);
    else
    {
        // rows == 2 < cols: Eigen preconditions the transpose; V's last column is the null vector of the
        // 2x3 system = normalised cross product of the two rows (up to sign, which cancels in F's error).
        std::vector<double> At(9, 0.0); // pad to 3x3 with a zero row: same right singular vectors
        for (int i = 0; i < rows; i++)
            for (int j = 0; j < 3; j++)
                At[(size_t)i + 3 * j] = A[(size_t)i + (size_t)rows * j];
        la::jacobi_svd_square(At.data(), 3, nullptr, S3, V3);
    }
    const double ep[3] = {V3[0 + 3 * 2], V3[1 + 3 * 2], V3[2 + 3 * 2]};
    // e_cross (row-major literal at :198) times H
    const double ex[9] = {0, ep[2], -ep[1], -ep[2], 0, ep[0], ep[1], -ep[0], 0}; // column-major of [e]_x
    double Fc[9];
    for (int cidx = 0; cidx < 3; cidx++)
        for (int r = 0; r < 3; r++)
            Fc[r + 3 * cidx] =
                (ex[r] * h.M[0 + 3 * cidx] + ex[r + 3] * h.M[1 + 3 * cidx]) + ex[r + 6] * h.M[2 + 3 * cidx];
    double U[9], S[3], V[9];
    la::jacobi_svd_square(Fc, 3, U, S, V);
    S[2] = 0;
    double Fcand[9];
    for (int cidx = 0; cidx < 3; cidx++)
        for (int r = 0; r < 3; r++)
        {
            const double t0 = (U[r + 0] * S[0]) * V[cidx + 0];
            const double t1 = (U[r + 3] * S[1]) * V[cidx + 3];
            const double t2 = (U[r + 6] * S[2]) * V[cidx + 6];
            Fcand[r + 3 * cidx] = (t0 + t1) + t2;
        }
    double oldF[9];
    std::memcpy(oldF, m.M, sizeof oldF);
    std::vector<bool> old_inliers = inliers;
    std::memcpy(m.M, Fcand, sizeof oldF);
    const double cand_score = evaluate(m, c, n, inliers);
    std::memcpy(m.M, oldF, sizeof oldF);
    const double orig_score = evaluate(m, c, n, old_inliers);
    if (cand_score > orig_score)
        std::memcpy(m.M, Fcand, sizeof oldF);
    else
        inliers = old_inliers;
}

// =================================================================================================
// src/model_inliers/ransac.cpp
// =================================================================================================

static double fast_pow_k(int k, double d)
{
    // ransac.cpp:32-51
    double t = d * d;
    if (k == 4)
        return t * t;
    if (k == 5)
        return t * t * d;
    t = t * t;
    return t * t;
}

namespace
{
// The sampling state of ransac.cpp:72-158, shared by ransac() and hypothesis_stream().
struct Sampler
{
    const Corr *c;
    size_t n;
    int k;
    bool has_quality = false;
    std::vector<size_t> sorted_idx;
    std::vector<size_t> eval_order;
    std::default_random_engine generator{42}; // :98
    size_t prosac_n;

    Sampler(const Corr *c_, size_t n_, int k_) : c(c_), n(n_), k(k_)
    {
        for (size_t i = 0; i < n; i++)
            if (c[i].quality != 0)
            {
                has_quality = true;
                break;
            }
        if (has_quality) // :83-90
        {
            sorted_idx.resize(n);
            std::iota(sorted_idx.begin(), sorted_idx.end(), 0);
            std::sort(sorted_idx.begin(), sorted_idx.end(),
                      [this](size_t a, size_t b) { return c[a].quality < c[b].quality; });
        }
        eval_order.resize(n);
        std::iota(eval_order.begin(), eval_order.end(), 0);
        prosac_n = has_quality ? (size_t)k : n; // :100
        std::shuffle(eval_order.begin(), eval_order.end(), generator); // :158 (first RNG consumer)
    }
    size_t map_idx(size_t i) const { return has_quality ? sorted_idx[i] : i; }
    void random_k_from_n(size_t pool, size_t *indices) // :104-127
    {
        std::uniform_int_distribution<size_t> dist(0, pool - 1);
        for (int j = 0; j < k; j++)
        {
            size_t candidate;
            bool unique;
            do
            {
                candidate = dist(generator);
                unique = true;
                for (int t = 0; t < j; t++)
                    if (indices[t] == map_idx(candidate))
                    {
                        unique = false;
                        break;
                    }
            } while (!unique);
            indices[j] = map_idx(candidate);
        }
    }
    void prosac_sample(size_t pool, size_t *indices) // :130-154
    {
        indices[0] = sorted_idx[pool - 1];
        std::uniform_int_distribution<size_t> dist(0, pool - 2);
        for (int j = 1; j < k; j++)
        {
            size_t candidate;
            bool unique;
            do
            {
                candidate = dist(generator);
                unique = true;
                for (int t = 0; t < j; t++)
                    if (indices[t] == sorted_idx[candidate])
                    {
                        unique = false;
                        break;
                    }
            } while (!unique);
            indices[j] = sorted_idx[candidate];
        }
    }
    void next(size_t i, size_t *indices) // :164-171
    {
        if (has_quality && prosac_n < n && i > 0 && i % 10 == 0)
            prosac_n++;
        if (has_quality && prosac_n < n && prosac_n > (size_t)k)
            prosac_sample(prosac_n, indices);
        else
            random_k_from_n(has_quality ? prosac_n : n, indices);
    }
};
} // namespace

void hypothesis_stream(const Corr *c, size_t n, int kind, size_t count, std::vector<size_t> &eval_order,
                       std::vector<size_t> &samples)
{
    const int k = minimum_points(kind);
    samples.clear();
    eval_order.clear();
    if (n < (size_t)k)
        return;
    Sampler s(c, n, k);
    eval_order = s.eval_order;
    samples.resize(count * k);
    for (size_t i = 0; i < count; i++)
        s.next(i, &samples[i * k]);
}

double ransac(const Corr *c, size_t n, Model &model, std::vector<bool> &inliers, RansacTrace *trace)
{
    const size_t MIN_ITERATIONS = 20;
    const size_t MAX_ITERATIONS = 10000;
    const size_t MAX_INNER_ITERATIONS = 5;
    const double PROBABILITY = 0.999;
    const double log_1m_p = std::log(1 - PROBABILITY);
    const int k = minimum_points(model.kind);
    RansacTrace tr;

    inliers.resize(n);
    std::fill(inliers.begin(), inliers.end(), false);
    if (n < (size_t)k)
    {
        if (trace)
            *trace = tr;
        return 0;
    }
    Sampler sampler(c, n, k);
    const std::vector<size_t> &eval_order = sampler.eval_order;

    Model best_model(model.kind);
    best_model.thr = model.thr;
    double best_score = 0;
    size_t probability_iterations = MAX_ITERATIONS;
    std::vector<bool> candidate_inliers(n, false);
    size_t sample[8];

    size_t i = 0;
    for (; i < probability_iterations; i++)
    {
        sampler.next(i, sample);
        if (model.kind == MODEL_HOMOGRAPHY && check_sample_degeneracy_h(c, sample)) // :173-177
        {
            tr.degenerate++;
            continue;
        }
        fit(model, c, sample); // :179

        double score = 0;
        size_t checked = 0;
        bool rejected = false;
        std::fill(candidate_inliers.begin(), candidate_inliers.end(), false);
        for (size_t idx : eval_order) // :187-203
        {
            const double e = error(model, c[idx]);
            if (e < model.thr)
            {
                candidate_inliers[idx] = true;
                const double ratio = e / model.thr;
                score += 1.0 - ratio * ratio;
            }
            checked++;
            if (checked > 20 && best_score > 0 && score < best_score * static_cast<double>(checked) / n * 0.6)
            {
                rejected = true;
                break;
            }
        }
        if (rejected)
        {
            tr.rejected++;
            continue;
        }
        if (score > best_score) // :207
        {
            tr.improvements++;
            best_model = model;
            best_score = score;
            inliers = candidate_inliers;
            if (model.kind == MODEL_FUNDAMENTAL) // :213-222
            {
                check_degeneracy_f(model, c, n, inliers);
                const double degen_score = evaluate(model, c, n, inliers);
                if (degen_score > best_score)
                {
                    best_model = model;
                    best_score = degen_score;
                }
            }
            fit_inliers(model, c, n, inliers); // :224
            double inlier_score = evaluate(model, c, n, inliers);
            if (inlier_score > best_score)
            {
                best_model = model;
                best_score = inlier_score;
                for (size_t j = 1; j < MAX_INNER_ITERATIONS; j++)
                {
                    fit_inliers(model, c, n, inliers);
                    inlier_score = evaluate(model, c, n, inliers);
                    if (inlier_score > best_score)
                    {
                        best_model = model;
                        best_score = inlier_score;
                    }
                    else
                    {
                        break;
                    }
                }
            }
            const double omega = best_score / n; // :247-251
            const double omega_n = fast_pow_k(k, omega);
            const double log_1m_omega_n = std::log(1 - omega_n);
            probability_iterations =
                std::max(MIN_ITERATIONS, std::min(MAX_ITERATIONS, static_cast<size_t>(log_1m_p / log_1m_omega_n)));
        }
    }
    tr.iterations = i;
    if (trace)
        *trace = tr;
    model = best_model;
    return evaluate(model, c, n, inliers) / n; // :255-256
}

} // namespace oc_oracle
