// Small dense linear algebra for the host side of the RANSAC path: the three Eigen 3.4 algorithms the
// reference's model fits call (reference src/model_inliers/homography_model.cpp:44,49,81,86;
// essential_matrix_model.cpp:16,22; fundamental_matrix_model.cpp:18,25,188,196), restated from their published
// descriptions because Eigen is an external, un-vendored dependency of the reference:
//   full-pivoting LU solve (rank-revealing, not least squares, also for tall systems), cofactor 3x3 inverse,
//   two-sided Jacobi SVD with a column-pivoting Householder QR preconditioner for tall matrices.
// Matrices are passed as ColMat (column-major, like Eigen's default) views.
#pragma once
#include <cstddef>
#include <vector>

namespace ocb_host
{
namespace linalg
{

struct ColMat
{
    int rows = 0, cols = 0;
    std::vector<double> a;
    ColMat() {}
    ColMat(int r, int c) : rows(r), cols(c), a((size_t)r * c, 0.0) {}
    double &operator()(int r, int c) { return a[(size_t)c * rows + r]; }
    double operator()(int r, int c) const { return a[(size_t)c * rows + r]; }
};

// x = FullPivLU(A).solve(b); x.size() == A.cols
std::vector<double> full_piv_lu_solve(const ColMat &A, const std::vector<double> &b);

// 3x3 inverse by cofactors; in/out column-major [9]
void invert3(const double *m, double *out);

struct Svd
{
    ColMat U, V;              // U: rows x rows (square input) ; V: cols x cols
    std::vector<double> sigma; // descending
};
// Jacobi SVD of a square matrix. want_u / want_v select which factors are accumulated.
Svd jacobi_svd(const ColMat &A, bool want_u, bool want_v);
// Right singular vectors + singular values of a tall (rows >= cols) matrix.
Svd jacobi_svd_tall(const ColMat &A);

} // namespace linalg
} // namespace ocb_host
