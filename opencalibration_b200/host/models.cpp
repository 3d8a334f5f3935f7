// Member functions of the reference's three RANSAC models on top of libocb.so.
//   reference src/model_inliers/homography_model.cpp, essential_matrix_model.cpp, fundamental_matrix_model.cpp
// Bulk scoring (evaluate) runs on the GPU through ocb_score_models; the minimal / all-inlier fits are tiny
// dense solves and stay on the host (linalg.hpp); error() is the reference's one-correspondence scalar accessor
// and is evaluated in the same canonical operation order the kernels use (no FMA: build with -ffp-contract=off).
#include "models_detail.hpp"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <limits>
#include <stdexcept>
#include <string>

namespace ocb_host
{
namespace detail
{

void gpu_check(int rc, const char *what)
{
    if (rc != 0)
        throw std::runtime_error(std::string(what) + " failed (" + std::to_string(rc) + "): " + ocb_last_error());
}

const double *corr_data(const std::vector<opencalibration::correspondence> &c)
{
    return reinterpret_cast<const double *>(c.data());
}

// measurement.hnormalized() -- (x/z, y/z), two true divisions
static inline void hnorm(const Eigen::Vector3d &m, double &x, double &y)
{
    x = m[0] / m[2];
    y = m[1] / m[2];
}

// measurement / measurement.z for the residuals: (x/z, y/z, z/z); when z/z is not exactly 1 (z is zero,
// infinite or NaN) every term of the reference residual is NaN -- same convention as the kernels.
static inline void unit_depth(const opencalibration::correspondence &c, double &x1, double &y1, double &x2, double &y2)
{
    hnorm(c.measurement1, x1, y1);
    hnorm(c.measurement2, x2, y2);
    const double z1 = c.measurement1[2], z2 = c.measurement2[2];
    if (!(z1 / z1 == 1.0))
        x1 = y1 = std::numeric_limits<double>::quiet_NaN();
    if (!(z2 / z2 == 1.0))
        x2 = y2 = std::numeric_limits<double>::quiet_NaN();
}

double homography_error(const double *H, const double *G, const opencalibration::correspondence &c)
{
    // homography_model.cpp:89-97
    double x1, y1, x2, y2;
    unit_depth(c, x1, y1, x2, y2);
    const double px = (H[0] * x1 + H[3] * y1) + H[6];
    const double py = (H[1] * x1 + H[4] * y1) + H[7];
    const double pz = (H[2] * x1 + H[5] * y1) + H[8];
    const double fx = px / pz - x2, fy = py / pz - y2;
    const double forward = fx * fx + fy * fy;
    const double qx = (G[0] * x2 + G[3] * y2) + G[6];
    const double qy = (G[1] * x2 + G[4] * y2) + G[7];
    const double qz = (G[2] * x2 + G[5] * y2) + G[8];
    const double bx = qx / qz - x1, by = qy / qz - y1;
    const double backward = bx * bx + by * by;
    return std::sqrt((forward + backward) / 2.0);
}

double epipolar_error(const double *E, const opencalibration::correspondence &c)
{
    // essential_matrix_model.cpp:112-123 == fundamental_matrix_model.cpp:110-121
    double x1, y1, x2, y2;
    unit_depth(c, x1, y1, x2, y2);
    const double l0 = (x2 * E[0] + y2 * E[1]) + E[2]; // x2^T E, also (E^T x2)
    const double l1 = (x2 * E[3] + y2 * E[4]) + E[5];
    const double l2 = (x2 * E[6] + y2 * E[7]) + E[8];
    const double num = (l0 * x1 + l1 * y1) + l2;
    const double m0 = (E[0] * x1 + E[3] * y1) + E[6]; // E x1
    const double m1 = (E[1] * x1 + E[4] * y1) + E[7];
    const double denom = ((m0 * m0 + m1 * m1) + l0 * l0) + l1 * l1;
    if (denom < 1e-20)
        return std::numeric_limits<double>::max();
    return std::sqrt((num * num) / denom);
}

namespace
{
struct BoundState
{
    const void *data = nullptr;
    size_t n = 0;
    const uint32_t *order = nullptr;
};
thread_local BoundState t_bound;
bool is_bound(const std::vector<opencalibration::correspondence> &corrs)
{
    return t_bound.data != nullptr && t_bound.data == static_cast<const void *>(corrs.data()) &&
           t_bound.n == corrs.size();
}
} // namespace

BoundCorrespondences::BoundCorrespondences(const std::vector<opencalibration::correspondence> &corrs,
                                           const uint32_t *order)
    : prev_data_(t_bound.data), prev_n_(t_bound.n), prev_order_(t_bound.order)
{
    t_bound = BoundState();
    if (corrs.empty())
        return;
    gpu_check(ocb_corr_bind(corr_data(corrs), corrs.size(), order), "ocb_corr_bind");
    t_bound.data = corrs.data(), t_bound.n = corrs.size(), t_bound.order = order;
}

BoundCorrespondences::~BoundCorrespondences()
{
    t_bound = BoundState();
    ocb_corr_unbind();
    if (prev_data_ && ocb_corr_bind(static_cast<const double *>(prev_data_), prev_n_, prev_order_) == 0)
        t_bound.data = prev_data_, t_bound.n = prev_n_, t_bound.order = prev_order_;
}

void gpu_evaluate_bits(int kind, const double *m18, double thr, const std::vector<opencalibration::correspondence> &corrs,
                       double *score, uint32_t *count, uint32_t *bits)
{
    if (is_bound(corrs))
        gpu_check(ocb_score_bound(kind, m18, 1, thr, 0, score, count, bits), "ocb_score_bound");
    else
        gpu_check(ocb_score_models(kind, m18, 1, corr_data(corrs), corrs.size(), thr, nullptr, score, count, bits),
                  "ocb_score_models");
}

void gpu_residuals(int kind, const double *m18, const std::vector<opencalibration::correspondence> &corrs, double *e)
{
    if (is_bound(corrs))
        gpu_check(ocb_residuals_bound(kind, m18, e), "ocb_residuals_bound");
    else
        gpu_check(ocb_residuals(kind, m18, corr_data(corrs), corrs.size(), e), "ocb_residuals");
}

void gpu_score_in_order(int kind, const double *models18, size_t h,
                        const std::vector<opencalibration::correspondence> &corrs, double thr, const uint32_t *order,
                        double *score, uint32_t *count)
{
    if (is_bound(corrs) && t_bound.order == order && order != nullptr)
        gpu_check(ocb_score_bound(kind, models18, h, thr, 1, score, count, nullptr), "ocb_score_bound");
    else
        gpu_check(ocb_score_models(kind, models18, h, corr_data(corrs), corrs.size(), thr, order, score, count, nullptr),
                  "ocb_score_models");
}

double gpu_evaluate(int kind, const double *matrix9, const double *inverse9, double thr,
                    const std::vector<opencalibration::correspondence> &corrs, std::vector<bool> &inliers)
{
    // Model::evaluate (homography_model.cpp:99-118 and twins): index-order MSAC sum + inlier flags
    const size_t n = corrs.size();
    inliers.resize(n);
    if (n == 0)
        return 0.0;
    double m18[18];
    std::memcpy(m18, matrix9, sizeof(double) * 9);
    if (inverse9)
        std::memcpy(m18 + 9, inverse9, sizeof(double) * 9);
    else
        std::fill(m18 + 9, m18 + 18, 0.0);
    std::vector<uint32_t> bits((n + 31) / 32);
    double score = 0.0;
    uint32_t count = 0;
    gpu_evaluate_bits(kind, m18, thr, corrs, &score, &count, bits.data());
    for (size_t i = 0; i < n; i++)
        inliers[i] = (bits[i >> 5] >> (i & 31)) & 1u;
    return score;
}

// ---- homography fit -------------------------------------------------------------------------------------
static void dlt_rows(const opencalibration::correspondence &c, linalg::ColMat &P, int row)
{
    // homography_model.cpp:26-35
    double x, y, u, v;
    hnorm(c.measurement1, x, y);
    hnorm(c.measurement2, u, v);
    const double top[9] = {-x, -y, -1, 0, 0, 0, x * u, y * u, u};
    const double bot[9] = {0, 0, 0, -x, -y, -1, x * v, y * v, v};
    for (int k = 0; k < 9; k++)
    {
        P(row, k) = top[k];
        P(row + 1, k) = bot[k];
    }
}

static void finish_homography(linalg::ColMat &P, opencalibration::homography_model &m)
{
    // last row: h33 == 1 (homography_model.cpp:37-49 / :76-86)
    const int last = P.rows - 1;
    for (int k = 0; k < 9; k++)
        P(last, k) = 0.0;
    P(last, 8) = 1.0;
    std::vector<double> rhs(P.rows, 0.0);
    rhs[last] = 1.0;
    const std::vector<double> h = linalg::full_piv_lu_solve(P, rhs);
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++)
            m.homography(r, c) = h[3 * r + c];
    const double h22 = m.homography(2, 2);
    double *H = m.homography.data();
    for (int i = 0; i < 9; i++)
        H[i] /= h22;
    linalg::invert3(H, m.homography_inverse.data());
}

void fit_homography(opencalibration::homography_model &m, const std::vector<opencalibration::correspondence> &corrs,
                    const size_t *sample, size_t count)
{
    linalg::ColMat P(int(2 * count + 1), 9);
    for (size_t i = 0; i < count; i++)
        dlt_rows(corrs[sample[i]], P, int(2 * i));
    finish_homography(P, m);
}

// ---- epipolar fits ----------------------------------------------------------------------------------------
static void epipolar_row(const opencalibration::correspondence &c, double *row)
{
    // essential_matrix_model.cpp:52-59
    double x, y, u, v;
    hnorm(c.measurement1, x, y);
    hnorm(c.measurement2, u, v);
    const double r[9] = {x * u, x * v, x, y * u, y * v, y, u, v, 1};
    std::memcpy(row, r, sizeof r);
}

static void recompose(const linalg::Svd &s, const double *sigma, double *out9)
{
    // (U * diag(sigma)) * V^T
    for (int c = 0; c < 3; c++)
        for (int r = 0; r < 3; r++)
        {
            const double a = (s.U(r, 0) * sigma[0]) * s.V(c, 0);
            const double b = (s.U(r, 1) * sigma[1]) * s.V(c, 1);
            const double d = (s.U(r, 2) * sigma[2]) * s.V(c, 2);
            out9[r + 3 * c] = (a + b) + d;
        }
}

void fit_epipolar(double *matrix9, bool essential, const std::vector<std::array<double, 9>> &rows)
{
    // calculateEssentialMatrix / calculateFundamentalMatrix
    // (essential_matrix_model.cpp:12-31, fundamental_matrix_model.cpp:13-29)
    linalg::ColMat AtA(9, 9);
    for (int j = 0; j < 9; j++)
        for (int i = 0; i < 9; i++)
        {
            double acc = 0.0;
            for (const auto &r : rows)
                acc += r[i] * r[j];
            AtA(i, j) = acc;
        }
    const linalg::Svd big = linalg::jacobi_svd(AtA, false, true);
    linalg::ColMat F(3, 3);
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++)
            F(r, c) = big.V(3 * r + c, 8);
    const linalg::Svd s = linalg::jacobi_svd(F, true, true);
    double sigma[3] = {s.sigma[0], s.sigma[1], 0.0};
    if (essential)
        sigma[0] = sigma[1] = (s.sigma[0] + s.sigma[1]) / 2.0;
    recompose(s, sigma, matrix9);
}

void epipolar_rows_from_sample(const std::vector<opencalibration::correspondence> &corrs, const size_t *sample,
                               size_t count, std::vector<std::array<double, 9>> &rows)
{
    rows.resize(count);
    for (size_t i = 0; i < count; i++)
        epipolar_row(corrs[sample[i]], rows[i].data());
}

bool epipolar_rows_from_inliers(const std::vector<opencalibration::correspondence> &corrs,
                                const std::vector<bool> &inliers, size_t minimum,
                                std::vector<std::array<double, 9>> &rows)
{
    const size_t num = (size_t)std::count(inliers.begin(), inliers.end(), true);
    if (num < minimum)
        return false; // essential_matrix_model.cpp:65-66
    rows.clear();
    rows.reserve(num);
    for (size_t i = 0; i < corrs.size(); i++)
        if (inliers[i])
        {
            rows.emplace_back();
            epipolar_row(corrs[i], rows.back().data());
        }
    return true;
}

void rank2_from(const double *in9, double *out9)
{
    linalg::ColMat F(3, 3);
    for (int i = 0; i < 9; i++)
        F.a[i] = in9[i];
    const linalg::Svd s = linalg::jacobi_svd(F, true, true);
    const double sigma[3] = {s.sigma[0], s.sigma[1], 0.0};
    recompose(s, sigma, out9);
}

} // namespace detail
} // namespace ocb_host

namespace opencalibration
{
using namespace ocb_host::detail;
namespace la = ocb_host::linalg;

// =============================================================================================================
// homography_model
// =============================================================================================================
homography_model::homography_model()
    : homography(Eigen::Matrix3d::Constant(NAN)), homography_inverse(Eigen::Matrix3d::Constant(NAN))
{
}

void homography_model::fit(const std::vector<correspondence> &corrs,
                           const std::array<size_t, MINIMUM_POINTS> &initial_indices)
{
    fit_homography(*this, corrs, initial_indices.data(), MINIMUM_POINTS);
}

void homography_model::fitInliers(const std::vector<correspondence> &corrs, const std::vector<bool> &inliers)
{
    std::vector<size_t> idx;
    for (size_t i = 0; i < corrs.size(); i++)
        if (inliers[i])
            idx.push_back(i);
    fit_homography(*this, corrs, idx.data(), idx.size());
}

double homography_model::error(const correspondence &cor)
{
    return homography_error(homography.data(), homography_inverse.data(), cor);
}

double homography_model::evaluate(const std::vector<correspondence> &corrs, std::vector<bool> &inliers)
{
    return gpu_evaluate(OCB_MODEL_HOMOGRAPHY, homography.data(), homography_inverse.data(), inlier_threshold, corrs,
                        inliers);
}

bool homography_model::checkSampleDegeneracy(const std::vector<correspondence> &corrs,
                                             const std::array<size_t, MINIMUM_POINTS> &indices)
{
    // homography_model.cpp:120-136: any three of the four source points (nearly) collinear
    double px[4], py[4];
    for (int i = 0; i < 4; i++)
    {
        const Eigen::Vector3d &m = corrs[indices[i]].measurement1;
        px[i] = m[0] / m[2];
        py[i] = m[1] / m[2];
    }
    for (int a = 0; a < 4; a++)
        for (int b = a + 1; b < 4; b++)
            for (int c = b + 1; c < 4; c++)
            {
                const double ux = px[b] - px[a], uy = py[b] - py[a];
                const double wx = px[c] - px[a], wy = py[c] - py[a];
                if (std::abs(ux * wy - uy * wx) < 1e-10)
                    return true;
            }
    return false;
}

// =============================================================================================================
// essential_matrix_model
// =============================================================================================================
essential_matrix_model::essential_matrix_model() : essential_matrix(Eigen::Matrix3d::Constant(NAN))
{
}

void essential_matrix_model::fit(const std::vector<correspondence> &corrs,
                                 const std::array<size_t, MINIMUM_POINTS> &initial_indices)
{
    std::vector<std::array<double, 9>> rows;
    epipolar_rows_from_sample(corrs, initial_indices.data(), MINIMUM_POINTS, rows);
    fit_epipolar(essential_matrix.data(), true, rows);
}

void essential_matrix_model::fitInliers(const std::vector<correspondence> &corrs, const std::vector<bool> &inliers)
{
    std::vector<std::array<double, 9>> rows;
    if (epipolar_rows_from_inliers(corrs, inliers, MINIMUM_POINTS, rows))
        fit_epipolar(essential_matrix.data(), true, rows);
}

double essential_matrix_model::error(const correspondence &cor)
{
    return epipolar_error(essential_matrix.data(), cor);
}

double essential_matrix_model::evaluate(const std::vector<correspondence> &corrs, std::vector<bool> &inliers)
{
    return gpu_evaluate(OCB_MODEL_ESSENTIAL, essential_matrix.data(), nullptr, inlier_threshold, corrs, inliers);
}

bool essential_matrix_model::decompose(const std::vector<correspondence> & /*corrs*/,
                                       const std::vector<bool> & /*inliers*/, std::array<decomposed_pose, 4> &poses)
{
    // essential_matrix_model.cpp:125-153: R = U W V^T / U W^T V^T (det fixed to +1), t = +-U.col(2)
    la::ColMat E(3, 3);
    for (int i = 0; i < 9; i++)
        E.a[i] = essential_matrix.data()[i];
    const la::Svd s = la::jacobi_svd(E, true, true);
    const double W[9] = {0, 1, 0, -1, 0, 0, 0, 0, 1}; // column-major of [[0,-1,0],[1,0,0],[0,0,1]]
    auto compose = [&](bool transposeW, double *R) {
        double UW[9];
        for (int c = 0; c < 3; c++)
            for (int r = 0; r < 3; r++)
            {
                double acc = 0;
                for (int k = 0; k < 3; k++)
                    acc += s.U(r, k) * (transposeW ? W[c + 3 * k] : W[k + 3 * c]);
                UW[r + 3 * c] = acc;
            }
        for (int c = 0; c < 3; c++)
            for (int r = 0; r < 3; r++)
            {
                double acc = 0;
                for (int k = 0; k < 3; k++)
                    acc += UW[r + 3 * k] * s.V(c, k);
                R[r + 3 * c] = acc;
            }
        const double det = R[0] * (R[4] * R[8] - R[7] * R[5]) - R[3] * (R[1] * R[8] - R[7] * R[2]) +
                           R[6] * (R[1] * R[5] - R[4] * R[2]);
        if (det < 0)
            for (int i = 0; i < 9; i++)
                R[i] = -R[i];
    };
    double R1[9], R2[9];
    compose(false, R1);
    compose(true, R2);
    const Eigen::Vector3d t(s.U(0, 2), s.U(1, 2), s.U(2, 2)), mt(-t[0], -t[1], -t[2]);
    const Eigen::Quaterniond q1 = quaternion_from_rotation(R1), q2 = quaternion_from_rotation(R2);
    poses[0].orientation = q1, poses[0].position = t;
    poses[1].orientation = q1, poses[1].position = mt;
    poses[2].orientation = q2, poses[2].position = t;
    poses[3].orientation = q2, poses[3].position = mt;
    return true;
}

// =============================================================================================================
// fundamental_matrix_model
// =============================================================================================================
fundamental_matrix_model::fundamental_matrix_model() : fundamental_matrix(Eigen::Matrix3d::Constant(NAN))
{
}

void fundamental_matrix_model::fit(const std::vector<correspondence> &corrs,
                                   const std::array<size_t, MINIMUM_POINTS> &initial_indices)
{
    std::vector<std::array<double, 9>> rows;
    epipolar_rows_from_sample(corrs, initial_indices.data(), MINIMUM_POINTS, rows);
    fit_epipolar(fundamental_matrix.data(), false, rows);
}

void fundamental_matrix_model::fitInliers(const std::vector<correspondence> &corrs, const std::vector<bool> &inliers)
{
    std::vector<std::array<double, 9>> rows;
    if (epipolar_rows_from_inliers(corrs, inliers, MINIMUM_POINTS, rows))
        fit_epipolar(fundamental_matrix.data(), false, rows);
}

double fundamental_matrix_model::error(const correspondence &cor)
{
    return epipolar_error(fundamental_matrix.data(), cor);
}

double fundamental_matrix_model::evaluate(const std::vector<correspondence> &corrs, std::vector<bool> &inliers)
{
    return gpu_evaluate(OCB_MODEL_FUNDAMENTAL, fundamental_matrix.data(), nullptr, inlier_threshold, corrs, inliers);
}

void fundamental_matrix_model::checkDegeneracy(const std::vector<correspondence> &corrs, std::vector<bool> &inliers)
{
    // DEGENSAC, fundamental_matrix_model.cpp:123-215: if the F-inliers are dominated by a plane,
    // rebuild F = [e']_x H from the plane homography and the off-plane points.
    std::vector<size_t> f_idx;
    for (size_t i = 0; i < inliers.size(); i++)
        if (inliers[i])
            f_idx.push_back(i);
    if (f_idx.size() < homography_model::MINIMUM_POINTS)
        return;

    homography_model plane;
    plane.inlier_threshold = inlier_threshold * 2;
    std::array<size_t, 4> seed{f_idx[0], f_idx[1], f_idx[2], f_idx[3]};
    plane.fit(corrs, seed);

    // residuals of the F-inliers under the plane homography, on the GPU (one model, all correspondences)
    std::vector<double> e(corrs.size());
    auto plane_residuals = [&]() {
        double m18[18];
        std::memcpy(m18, plane.homography.data(), 72);
        std::memcpy(m18 + 9, plane.homography_inverse.data(), 72);
        gpu_residuals(OCB_MODEL_HOMOGRAPHY, m18, corrs, e.data());
    };
    plane_residuals();
    std::vector<bool> on_plane(corrs.size(), false);
    size_t plane_count = 0;
    for (size_t idx : f_idx)
        if (e[idx] < plane.inlier_threshold)
        {
            on_plane[idx] = true;
            plane_count++;
        }
    if (static_cast<double>(plane_count) / f_idx.size() < 0.7)
        return;

    plane.fitInliers(corrs, on_plane);
    plane_residuals();
    std::vector<size_t> off_plane;
    for (size_t idx : f_idx)
    {
        if (e[idx] < plane.inlier_threshold)
            on_plane[idx] = true;
        else
        {
            on_plane[idx] = false;
            off_plane.push_back(idx);
        }
    }
    if (off_plane.size() < 2)
        return;

    // epipole: (x2 x H x1) . e' = 0 for every off-plane point
    const double *H = plane.homography.data();
    la::ColMat A(int(off_plane.size()), 3);
    for (size_t i = 0; i < off_plane.size(); i++)
    {
        const correspondence &c = corrs[off_plane[i]];
        const double a[3] = {c.measurement1[0] / c.measurement1[2], c.measurement1[1] / c.measurement1[2],
                             c.measurement1[2] / c.measurement1[2]};
        const double b[3] = {c.measurement2[0] / c.measurement2[2], c.measurement2[1] / c.measurement2[2],
                             c.measurement2[2] / c.measurement2[2]};
        double ha[3];
        for (int r = 0; r < 3; r++)
            ha[r] = (H[r] * a[0] + H[r + 3] * a[1]) + H[r + 6] * a[2];
        A(int(i), 0) = b[1] * ha[2] - b[2] * ha[1];
        A(int(i), 1) = b[2] * ha[0] - b[0] * ha[2];
        A(int(i), 2) = b[0] * ha[1] - b[1] * ha[0];
    }
    la::Svd sv;
    if (A.rows >= 3)
        sv = la::jacobi_svd_tall(A);
    else
    {
        la::ColMat padded(3, 3); // two equations: pad with a zero row, same null space
        for (int r = 0; r < A.rows; r++)
            for (int c = 0; c < 3; c++)
                padded(r, c) = A(r, c);
        sv = la::jacobi_svd(padded, false, true);
    }
    const double ep[3] = {sv.V(0, 2), sv.V(1, 2), sv.V(2, 2)};
    const double cross[9] = {0, ep[2], -ep[1], -ep[2], 0, ep[0], ep[1], -ep[0], 0}; // [e']_x, column-major
    double candidate[9];
    for (int c = 0; c < 3; c++)
        for (int r = 0; r < 3; r++)
            candidate[r + 3 * c] = (cross[r] * H[0 + 3 * c] + cross[r + 3] * H[1 + 3 * c]) + cross[r + 6] * H[2 + 3 * c];
    double rank2[9];
    rank2_from(candidate, rank2);

    // keep the candidate only if it scores better (:201-214)
    const Eigen::Matrix3d old_F = fundamental_matrix;
    std::vector<bool> old_inliers = inliers;
    std::memcpy(fundamental_matrix.data(), rank2, sizeof rank2);
    const double candidate_score = evaluate(corrs, inliers);
    fundamental_matrix = old_F;
    const double original_score = evaluate(corrs, old_inliers);
    if (candidate_score > original_score)
        std::memcpy(fundamental_matrix.data(), rank2, sizeof rank2);
    else
        inliers = old_inliers;
}

} // namespace opencalibration
