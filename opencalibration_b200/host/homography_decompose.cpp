// homography_model::decompose with the reference's signature and post-processing
// (reference src/model_inliers/homography_model.cpp:138-185).
//
// The reference delegates the algebra to cv::decomposeHomographyMat(H, I, Rs, Ts, Ns) (:146). OpenCV is an
// external, un-vendored dependency of the reference (find_package(OpenCV), CMakeLists.txt:40), so the analytical
// method it implements -- Malis & Vargas, "Deeper understanding of the homography decomposition for vision-based
// control" (INRIA RR-6303), the closed form built on S = H^T H - I -- is restated here for K = I:
//   1. scale H so that its middle singular value is 1;
//   2. S = H^T H - I; if ||S||_inf < 0.001 the homography is a pure rotation: one solution (R = H, t = n = 0);
//   3. otherwise the two plane normals follow from the opposites of the minors of S, the translations from
//      ||t||, rho and the sign of the pivot S_ii, and R = H (I - (2/v) t* n^T), flipped to det R > 0;
//      solutions come in the order (Ra, ta, na), (Ra, -ta, -na), (Rb, tb, nb), (Rb, -tb, -nb).
// Golden vectors produced by the real cv2.decomposeHomographyMat (tests/golden/make_decompose_vectors.py) pin this
// restatement to 1e-9; it cannot be bit-exact because step 1 uses OpenCV's own SVD.
//
// What follows the call is the reference's code path verbatim in behaviour: the cheirality vote over the inlier
// correspondences (:160-172), Eigen::Quaterniond(R) (:173), score -1 for unused slots (:176-179) and the
// std::stable_sort with the reference's own (non-strict) comparator (:180-181), executed by the same libstdc++.
#include "models_detail.hpp"

#include <algorithm>
#include <cmath>

namespace
{
namespace la = ocb_host::linalg;

struct PlaneMotion
{
    double R[9]; // column-major
    double t[3];
    double n[3];
};

inline double at(const double *M, int r, int c) // column-major 3x3
{
    return M[r + 3 * c];
}

// -(minor of M at (row, col)), written as the difference of the two products like the published closed form
double opposite_of_minor(const double *M, int row, int col)
{
    const int x1 = col == 0 ? 1 : 0, x2 = col == 2 ? 1 : 2;
    const int y1 = row == 0 ? 1 : 0, y2 = row == 2 ? 1 : 2;
    return at(M, y1, x2) * at(M, y2, x1) - at(M, y1, x1) * at(M, y2, x2);
}

inline int sign_of(double x)
{
    return x >= 0 ? 1 : -1;
}

double det3(const double *M)
{
    return at(M, 0, 0) * (at(M, 1, 1) * at(M, 2, 2) - at(M, 1, 2) * at(M, 2, 1)) -
           at(M, 0, 1) * (at(M, 1, 0) * at(M, 2, 2) - at(M, 1, 2) * at(M, 2, 0)) +
           at(M, 0, 2) * (at(M, 1, 0) * at(M, 2, 1) - at(M, 1, 1) * at(M, 2, 0));
}

// R = Hn (I - (2/v) t* n^T), sign fixed so that det R > 0
void rotation_from_tstar_n(const double *Hn, const double *tstar, const double *n, double v, double *R)
{
    double A[9];
    for (int c = 0; c < 3; c++)
        for (int r = 0; r < 3; r++)
            A[r + 3 * c] = (r == c ? 1.0 : 0.0) - (2.0 / v) * tstar[r] * n[c];
    for (int c = 0; c < 3; c++)
        for (int r = 0; r < 3; r++)
        {
            double acc = 0;
            for (int k = 0; k < 3; k++)
                acc += at(Hn, r, k) * A[k + 3 * c];
            R[r + 3 * c] = acc;
        }
    if (det3(R) < 0)
        for (int i = 0; i < 9; i++)
            R[i] = -R[i];
}

// decomposeHomographyMat(H, K = I): returns the number of solutions (1 or 4)
int decompose_homography(const double *H, PlaneMotion *out)
{
    la::ColMat Hm(3, 3);
    for (int i = 0; i < 9; i++)
        Hm.a[i] = H[i];
    const la::Svd svd = la::jacobi_svd(Hm, false, false);
    double Hn[9];
    const double scale = 1.0 / svd.sigma[1];
    for (int i = 0; i < 9; i++)
        Hn[i] = H[i] * scale;

    double S[9];
    for (int c = 0; c < 3; c++)
        for (int r = 0; r < 3; r++)
        {
            double acc = 0;
            for (int k = 0; k < 3; k++)
                acc += at(Hn, k, r) * at(Hn, k, c);
            S[r + 3 * c] = acc - (r == c ? 1.0 : 0.0);
        }
    double norm_inf = 0;
    for (int r = 0; r < 3; r++)
        norm_inf = std::max(norm_inf, std::abs(at(S, r, 0)) + std::abs(at(S, r, 1)) + std::abs(at(S, r, 2)));
    if (norm_inf < 0.001)
    {
        for (int i = 0; i < 9; i++)
            out[0].R[i] = Hn[i];
        for (int i = 0; i < 3; i++)
            out[0].t[i] = out[0].n[i] = 0.0;
        return 1;
    }

    const double M00 = opposite_of_minor(S, 0, 0), M11 = opposite_of_minor(S, 1, 1), M22 = opposite_of_minor(S, 2, 2);
    // The principal minors are non-negative in exact arithmetic; when one is exactly zero (e.g. the plane normal is a
    // coordinate axis, as in the reference's own unit tests) rounding can leave -1e-17, whose square root would turn
    // every solution into NaN. Clamp instead of propagating the NaN OpenCV would produce there.
    const double rtM00 = std::sqrt(std::max(M00, 0.0)), rtM11 = std::sqrt(std::max(M11, 0.0)),
                 rtM22 = std::sqrt(std::max(M22, 0.0));
    const double M01 = opposite_of_minor(S, 0, 1), M12 = opposite_of_minor(S, 1, 2), M02 = opposite_of_minor(S, 0, 2);
    const int e12 = sign_of(M12), e02 = sign_of(M02), e01 = sign_of(M01);
    const double nS00 = std::abs(at(S, 0, 0)), nS11 = std::abs(at(S, 1, 1)), nS22 = std::abs(at(S, 2, 2));
    int pivot = 0; // argmax |S_ii|
    if (nS00 < nS11)
    {
        pivot = 1;
        if (nS11 < nS22)
            pivot = 2;
    }
    else if (nS00 < nS22)
        pivot = 2;

    double npa[3], npb[3];
    switch (pivot)
    {
    case 0:
        npa[0] = at(S, 0, 0), npb[0] = at(S, 0, 0);
        npa[1] = at(S, 0, 1) + rtM22, npb[1] = at(S, 0, 1) - rtM22;
        npa[2] = at(S, 0, 2) + e12 * rtM11, npb[2] = at(S, 0, 2) - e12 * rtM11;
        break;
    case 1:
        npa[0] = at(S, 0, 1) + rtM22, npb[0] = at(S, 0, 1) - rtM22;
        npa[1] = at(S, 1, 1), npb[1] = at(S, 1, 1);
        npa[2] = at(S, 1, 2) - e02 * rtM00, npb[2] = at(S, 1, 2) + e02 * rtM00;
        break;
    default:
        npa[0] = at(S, 0, 2) + e01 * rtM11, npb[0] = at(S, 0, 2) - e01 * rtM11;
        npa[1] = at(S, 1, 2) + rtM00, npb[1] = at(S, 1, 2) - rtM00;
        npa[2] = at(S, 2, 2), npb[2] = at(S, 2, 2);
        break;
    }
    const double traceS = at(S, 0, 0) + at(S, 1, 1) + at(S, 2, 2);
    const double v = 2.0 * std::sqrt(std::max(1 + traceS - M00 - M11 - M22, 0.0));
    const double ESii = sign_of(at(S, pivot, pivot));
    const double r = std::sqrt(2 + traceS + v);
    const double n_t = std::sqrt(std::max(2 + traceS - v, 0.0));
    const double la_norm = std::sqrt(npa[0] * npa[0] + npa[1] * npa[1] + npa[2] * npa[2]);
    const double lb_norm = std::sqrt(npb[0] * npb[0] + npb[1] * npb[1] + npb[2] * npb[2]);
    double na[3], nb[3], ta_star[3], tb_star[3];
    for (int i = 0; i < 3; i++)
        na[i] = npa[i] / la_norm, nb[i] = npb[i] / lb_norm;
    const double half_nt = 0.5 * n_t, esii_t_r = ESii * r;
    for (int i = 0; i < 3; i++)
    {
        ta_star[i] = half_nt * (esii_t_r * nb[i] - n_t * na[i]);
        tb_star[i] = half_nt * (esii_t_r * na[i] - n_t * nb[i]);
    }
    double Ra[9], Rb[9], ta[3], tb[3];
    rotation_from_tstar_n(Hn, ta_star, na, v, Ra);
    rotation_from_tstar_n(Hn, tb_star, nb, v, Rb);
    for (int rr = 0; rr < 3; rr++)
    {
        ta[rr] = at(Ra, rr, 0) * ta_star[0] + at(Ra, rr, 1) * ta_star[1] + at(Ra, rr, 2) * ta_star[2];
        tb[rr] = at(Rb, rr, 0) * tb_star[0] + at(Rb, rr, 1) * tb_star[1] + at(Rb, rr, 2) * tb_star[2];
    }
    for (int s = 0; s < 4; s++)
    {
        const double *R = s < 2 ? Ra : Rb, *t = s < 2 ? ta : tb, *n = s < 2 ? na : nb;
        const double sg = (s & 1) ? -1.0 : 1.0;
        for (int i = 0; i < 9; i++)
            out[s].R[i] = R[i];
        for (int i = 0; i < 3; i++)
            out[s].t[i] = sg * t[i], out[s].n[i] = sg * n[i];
    }
    return 4;
}
} // namespace

namespace ocb_host
{
namespace detail
{
int decompose_homography_mat(const double *H9, double *R36, double *t12, double *n12)
{
    PlaneMotion m[4];
    const int k = decompose_homography(H9, m);
    for (int s = 0; s < k; s++)
    {
        for (int i = 0; i < 9; i++)
            R36[9 * s + i] = m[s].R[i];
        for (int i = 0; i < 3; i++)
            t12[3 * s + i] = m[s].t[i], n12[3 * s + i] = m[s].n[i];
    }
    return k;
}
} // namespace detail
} // namespace ocb_host

namespace opencalibration
{
bool homography_model::decompose(const std::vector<correspondence> &corrs, const std::vector<bool> &inliers,
                                 std::array<decomposed_pose, 4> &poses)
{
    PlaneMotion sol[4];
    const size_t solutions = (size_t)decompose_homography(homography.data(), sol);
    for (size_t i = 0; i < solutions; i++)
    {
        const double *R = sol[i].R, *N = sol[i].n;
        const double RN[3] = {at(R, 0, 0) * N[0] + at(R, 0, 1) * N[1] + at(R, 0, 2) * N[2],
                              at(R, 1, 0) * N[0] + at(R, 1, 1) * N[1] + at(R, 1, 2) * N[2],
                              at(R, 2, 0) * N[0] + at(R, 2, 1) * N[1] + at(R, 2, 2) * N[2]};
        poses[i].score = 0;
        for (size_t j = 0; j < corrs.size(); j++)
        {
            if (!inliers[j])
                continue;
            const auto &m1 = corrs[j].measurement1;
            const auto &m2 = corrs[j].measurement2;
            const double dot1 = N[0] * m1[0] + N[1] * m1[1] + N[2] * m1[2];
            const double dot2 = RN[0] * m2[0] + RN[1] * m2[1] + RN[2] * m2[2];
            if (dot1 >= 0 && dot2 >= 0)
                poses[i].score++;
        }
        poses[i].orientation = ocb_host::detail::quaternion_from_rotation(R);
        poses[i].position = Eigen::Vector3d(sol[i].t[0], sol[i].t[1], sol[i].t[2]);
    }
    for (size_t i = solutions; i < poses.size(); i++)
        poses[i].score = -1;
    std::stable_sort(poses.begin(), poses.end(),
                     [](const decomposed_pose &p1, const decomposed_pose &p2) { return p1.score >= p2.score; });
    return poses[0].score > 0;
}
} // namespace opencalibration
