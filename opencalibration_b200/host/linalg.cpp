// See linalg.hpp. Restated from the published Eigen 3.4 algorithms (FullPivLU, InverseImpl size 3, JacobiSVD);
// evaluation order of the individual updates is the natural column-oriented one, which real Eigen may
// vectorise differently -- fits therefore agree with a reference build to rounding, not bit for bit.
#include "linalg.hpp"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <numeric>

namespace ocb_host
{
namespace linalg
{

// The tall systems of fitInliers ((2n+1) x 9, n up to thousands) make this the hottest host routine of the RANSAC
// tail: the pivot search and the rank-1 update are plain column sweeps, cloned for wider vector units (the build is
// generic x86-64; -ffp-contract=off keeps every multiply and subtract individually rounded in all clones).
#if defined(__x86_64__) && defined(__GNUC__) && !defined(__clang__)
__attribute__((target_clones("avx512f", "avx2", "default")))
#endif
std::vector<double> full_piv_lu_solve(const ColMat &A, const std::vector<double> &b)
{
    const int R = A.rows, Cn = A.cols, K = std::min(R, Cn);
    ColMat lu = A;
    std::vector<int> row_swap(K), col_swap(K);
    int pivots = K;
    double biggest_pivot = 0.0;

    for (int k = 0; k < K; ++k)
    {
        // complete pivoting: largest magnitude of the trailing block, scanned column by column, first hit wins.
        // Two passes per column (a branch-free max reduction the compiler vectorises, then the first row holding
        // it) select the same element as a single scan with a strict comparison.
        int pr = k, pc = k;
        double best = std::fabs(lu(k, k));
        for (int c = k; c < Cn; ++c)
        {
            const double *col = &lu.a[(size_t)c * R];
            // eight independent running maxima: a single one is a serial dependency chain the compiler may not
            // reorder; the maximum itself does not depend on the order (NaNs lose every comparison either way)
            double m[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            int r = k;
            for (; r + 8 <= R; r += 8)
                for (int j = 0; j < 8; ++j)
                    m[j] = std::max(m[j], std::fabs(col[r + j]));
            for (; r < R; ++r)
                m[0] = std::max(m[0], std::fabs(col[r]));
            double cmax = 0.0;
            for (int j = 0; j < 8; ++j)
                cmax = std::max(cmax, m[j]);
            if (cmax > best)
            {
                best = cmax, pc = c;
                for (int r = k; r < R; ++r)
                    if (std::fabs(col[r]) == cmax)
                    {
                        pr = r;
                        break;
                    }
            }
        }
        if (best == 0.0)
        {
            pivots = k;
            for (int i = k; i < K; ++i)
                row_swap[i] = col_swap[i] = i;
            break;
        }
        biggest_pivot = std::max(biggest_pivot, best);
        row_swap[k] = pr;
        col_swap[k] = pc;
        if (pr != k)
            for (int c = 0; c < Cn; ++c)
                std::swap(lu(k, c), lu(pr, c));
        if (pc != k)
            for (int r = 0; r < R; ++r)
                std::swap(lu(r, k), lu(r, pc));
        const double d = lu(k, k);
        for (int r = k + 1; r < R; ++r)
            lu(r, k) /= d;
        const double *lk = &lu.a[(size_t)k * R];
        for (int c = k + 1; c < Cn; ++c)
        {
            const double top = lu(k, c);
            double *lc = &lu.a[(size_t)c * R];
            for (int r = k + 1; r < R; ++r)
                lc[r] -= lk[r] * top;
        }
    }

    // numerical rank with Eigen's default threshold eps * min(rows, cols)
    const double cut = std::fabs(biggest_pivot) * (DBL_EPSILON * static_cast<double>(K));
    int rank = 0;
    for (int i = 0; i < pivots; ++i)
        if (std::fabs(lu(i, i)) > cut)
            ++rank;

    std::vector<double> x(Cn, 0.0);
    if (rank == 0)
        return x;

    std::vector<double> y = b; // P b
    for (int k = 0; k < K; ++k)
        std::swap(y[k], y[row_swap[k]]);
    for (int c = 0; c < K; ++c) // forward substitution with the unit lower factor
    {
        const double yc = y[c];
        if (yc == 0.0)
            continue;
        for (int r = c + 1; r < K; ++r)
            y[r] -= yc * lu(r, c);
    }
    for (int r = Cn; r < R; ++r) // rows below the square part (tall systems)
    {
        double dot = 0.0;
        for (int c = 0; c < Cn; ++c)
            dot += lu(r, c) * y[c];
        y[r] -= dot;
    }
    for (int c = rank - 1; c >= 0; --c) // back substitution on the rank x rank upper block
    {
        if (y[c] == 0.0)
            continue;
        y[c] /= lu(c, c);
        const double yc = y[c];
        for (int r = 0; r < c; ++r)
            y[r] -= yc * lu(r, c);
    }
    for (int i = 0; i < rank; ++i)
        x[i] = y[i];
    for (int k = K - 1; k >= 0; --k) // undo the column permutation
        std::swap(x[k], x[col_swap[k]]);
    return x;
}

void invert3(const double *m, double *out)
{
    auto at = [m](int r, int c) { return m[r + 3 * c]; };
    auto minor = [&](int i, int j) {
        const int r1 = (i + 1) % 3, r2 = (i + 2) % 3, c1 = (j + 1) % 3, c2 = (j + 2) % 3;
        return at(r1, c1) * at(r2, c2) - at(r1, c2) * at(r2, c1);
    };
    const double k00 = minor(0, 0), k10 = minor(1, 0), k20 = minor(2, 0);
    const double det = (k00 * at(0, 0) + k10 * at(1, 0)) + k20 * at(2, 0);
    const double s = 1.0 / det;
    // inverse(r, c) = cofactor(c, r) / det
    out[0 + 3 * 0] = k00 * s;
    out[0 + 3 * 1] = k10 * s;
    out[0 + 3 * 2] = k20 * s;
    for (int r = 1; r < 3; ++r)
        for (int c = 0; c < 3; ++c)
            out[r + 3 * c] = minor(c, r) * s;
}

namespace
{
struct Givens
{
    double c = 1.0, s = 0.0;
    Givens adjoint() const { return Givens{c, -s}; }
    Givens then(const Givens &o) const { return Givens{c * o.c - s * o.s, c * o.s + s * o.c}; }
    bool identity() const { return c == 1.0 && s == 0.0; }
};

// (x, y) <- (c x + s y, -s x + c y) element-wise over two strided vectors
void rotate(double *x, size_t sx, double *y, size_t sy, int n, const Givens &g)
{
    if (g.identity())
        return;
    for (int i = 0; i < n; ++i, x += sx, y += sy)
    {
        const double a = *x, b = *y;
        *x = g.c * a + g.s * b;
        *y = -g.s * a + g.c * b;
    }
}

// symmetric 2x2 [[x, y], [y, z]] -> rotation that diagonalises it
Givens symmetric_jacobi(double x, double y, double z)
{
    Givens g;
    const double twice = 2.0 * std::fabs(y);
    if (twice < DBL_MIN)
        return g;
    const double tau = (x - z) / twice;
    const double w = std::sqrt(tau * tau + 1.0);
    const double t = tau > 0.0 ? 1.0 / (tau + w) : 1.0 / (tau - w);
    const double n = 1.0 / std::sqrt(t * t + 1.0);
    g.s = -(t > 0.0 ? 1.0 : -1.0) * (y / std::fabs(y)) * std::fabs(t) * n;
    g.c = n;
    return g;
}

// general real 2x2 -> (left, right) rotations with left * M * right diagonal
void two_sided(double a, double b, double c, double d, Givens &left, Givens &right)
{
    Givens sym; // first make the block symmetric
    const double t = a + d, diff = c - b;
    if (std::fabs(diff) >= DBL_MIN)
    {
        const double u = t / diff;
        const double h = std::sqrt(1.0 + u * u);
        sym.s = 1.0 / h;
        sym.c = u / h;
    }
    const double a2 = sym.c * a + sym.s * c, b2 = sym.c * b + sym.s * d;
    const double d2 = -sym.s * b + sym.c * d;
    right = symmetric_jacobi(a2, b2, d2);
    left = sym.then(right.adjoint());
}
} // namespace

Svd jacobi_svd(const ColMat &A, bool want_u, bool want_v)
{
    const int n = A.rows;
    Svd out;
    out.sigma.assign(n, 0.0);
    ColMat W = A;
    double scale = 0.0;
    for (double v : W.a)
        scale = std::max(scale, std::fabs(v));
    if (!(scale > 0.0) || !std::isfinite(scale))
        scale = 1.0;
    for (double &v : W.a)
        v /= scale;
    if (want_u)
    {
        out.U = ColMat(n, n);
        for (int i = 0; i < n; ++i)
            out.U(i, i) = 1.0;
    }
    if (want_v)
    {
        out.V = ColMat(n, n);
        for (int i = 0; i < n; ++i)
            out.V(i, i) = 1.0;
    }
    const double rel = 2.0 * DBL_EPSILON;
    double diag_max = 0.0;
    for (int i = 0; i < n; ++i)
        diag_max = std::max(diag_max, std::fabs(W(i, i)));

    for (int sweep = 0, dirty = 1; dirty && sweep < 1000; ++sweep)
    {
        dirty = 0;
        for (int p = 1; p < n; ++p)
            for (int q = 0; q < p; ++q)
            {
                const double tol = std::max(DBL_MIN, rel * diag_max);
                if (!(std::fabs(W(p, q)) > tol || std::fabs(W(q, p)) > tol))
                    continue;
                dirty = 1;
                Givens L, Rg;
                two_sided(W(p, p), W(p, q), W(q, p), W(q, q), L, Rg);
                rotate(&W.a[p], n, &W.a[q], n, n, L); // rows p, q
                if (want_u)
                    rotate(&out.U.a[(size_t)p * n], 1, &out.U.a[(size_t)q * n], 1, n, L);
                rotate(&W.a[(size_t)p * n], 1, &W.a[(size_t)q * n], 1, n, Rg.adjoint()); // columns p, q
                if (want_v)
                    rotate(&out.V.a[(size_t)p * n], 1, &out.V.a[(size_t)q * n], 1, n, Rg.adjoint());
                diag_max = std::max(diag_max, std::max(std::fabs(W(p, p)), std::fabs(W(q, q))));
            }
    }
    for (int i = 0; i < n; ++i)
    {
        const double d = W(i, i), mag = std::fabs(d);
        out.sigma[i] = mag;
        if (want_u && mag != 0.0)
        {
            const double sign = d / mag;
            for (int r = 0; r < n; ++r)
                out.U(r, i) *= sign;
        }
    }
    for (double &s : out.sigma)
        s *= scale;
    for (int i = 0; i < n; ++i) // selection sort, descending, carrying the columns along
    {
        int arg = i;
        for (int k = i + 1; k < n; ++k)
            if (out.sigma[k] > out.sigma[arg])
                arg = k;
        if (out.sigma[arg] == 0.0)
            break;
        if (arg == i)
            continue;
        std::swap(out.sigma[i], out.sigma[arg]);
        for (int r = 0; r < n; ++r)
        {
            if (want_u)
                std::swap(out.U(r, i), out.U(r, arg));
            if (want_v)
                std::swap(out.V(r, i), out.V(r, arg));
        }
    }
    return out;
}

Svd jacobi_svd_tall(const ColMat &A)
{
    const int R = A.rows, Cn = A.cols;
    ColMat Q = A; // reduced in place to the triangular factor
    std::vector<int> order(Cn);
    std::iota(order.begin(), order.end(), 0);
    for (int k = 0; k < std::min(R, Cn); ++k)
    {
        int arg = k;
        double big = -1.0;
        for (int c = k; c < Cn; ++c)
        {
            double nn = 0.0;
            for (int r = k; r < R; ++r)
                nn += Q(r, c) * Q(r, c);
            if (nn > big)
                big = nn, arg = c;
        }
        if (arg != k)
        {
            for (int r = 0; r < R; ++r)
                std::swap(Q(r, k), Q(r, arg));
            std::swap(order[k], order[arg]);
        }
        double below = 0.0;
        for (int r = k + 1; r < R; ++r)
            below += Q(r, k) * Q(r, k);
        if (below <= DBL_MIN)
            continue;
        const double head = Q(k, k);
        const double beta = head >= 0.0 ? -std::sqrt(head * head + below) : std::sqrt(head * head + below);
        const double tau = (beta - head) / beta;
        std::vector<double> v(R - k, 1.0);
        for (int r = k + 1; r < R; ++r)
            v[r - k] = Q(r, k) / (head - beta);
        for (int c = k; c < Cn; ++c)
        {
            double dot = 0.0;
            for (int r = k; r < R; ++r)
                dot += v[r - k] * Q(r, c);
            dot *= tau;
            for (int r = k; r < R; ++r)
                Q(r, c) -= dot * v[r - k];
        }
    }
    ColMat Rm(Cn, Cn);
    for (int c = 0; c < Cn; ++c)
        for (int r = 0; r <= c && r < R; ++r)
            Rm(r, c) = Q(r, c);
    Svd inner = jacobi_svd(Rm, false, true);
    Svd out;
    out.sigma = inner.sigma;
    out.V = ColMat(Cn, Cn);
    for (int c = 0; c < Cn; ++c)
        for (int r = 0; r < Cn; ++r)
            out.V(order[r], c) = inner.V(r, c);
    return out;
}

} // namespace linalg
} // namespace ocb_host
