// Internal helpers shared by models.cpp, ransac.cpp and match_features.cpp (host side of the path).
#pragma once
#include "linalg.hpp"
#include "opencalibration_api.hpp"

#include <ocb.h>

#include <array>
#include <vector>

namespace ocb_host
{
namespace detail
{
// throws std::runtime_error with ocb_last_error() when rc != 0 (the product has no CPU fallback)
void gpu_check(int rc, const char *what);
// std::vector<correspondence>::data() viewed as the [n][7] doubles the C ABI takes
const double *corr_data(const std::vector<opencalibration::correspondence> &c);

double homography_error(const double *H, const double *Hinv, const opencalibration::correspondence &c);
double epipolar_error(const double *E, const opencalibration::correspondence &c);
// While alive, the correspondences are resident on the GPU for the calling thread (ocb_corr_bind): every
// gpu_evaluate / gpu_residuals / gpu_score_in_order on the SAME vector skips the upload. ransac() holds one for
// the duration of a run. Nesting restores the outer binding on destruction.
class BoundCorrespondences
{
  public:
    BoundCorrespondences(const std::vector<opencalibration::correspondence> &corrs, const uint32_t *order);
    ~BoundCorrespondences();
    BoundCorrespondences(const BoundCorrespondences &) = delete;
    BoundCorrespondences &operator=(const BoundCorrespondences &) = delete;

  private:
    const void *prev_data_;
    size_t prev_n_;
    const uint32_t *prev_order_;
};
// Model::evaluate on the GPU for one packed model: index-order MSAC sum, inlier count and bit mask
void gpu_evaluate_bits(int kind, const double *m18, double thr, const std::vector<opencalibration::correspondence> &corrs,
                       double *score, uint32_t *count, uint32_t *bits);
// Model::error for every correspondence (index order) of one model
void gpu_residuals(int kind, const double *m18, const std::vector<opencalibration::correspondence> &corrs, double *e);
// the score loop of ransac.cpp:183-196 for a batch of models, summed in `order`
void gpu_score_in_order(int kind, const double *models18, size_t h,
                        const std::vector<opencalibration::correspondence> &corrs, double thr, const uint32_t *order,
                        double *score, uint32_t *count);
double gpu_evaluate(int kind, const double *matrix9, const double *inverse9, double thr,
                    const std::vector<opencalibration::correspondence> &corrs, std::vector<bool> &inliers);

void fit_homography(opencalibration::homography_model &m, const std::vector<opencalibration::correspondence> &corrs,
                    const size_t *sample, size_t count);
void fit_epipolar(double *matrix9, bool essential, const std::vector<std::array<double, 9>> &rows);
void epipolar_rows_from_sample(const std::vector<opencalibration::correspondence> &corrs, const size_t *sample,
                               size_t count, std::vector<std::array<double, 9>> &rows);
bool epipolar_rows_from_inliers(const std::vector<opencalibration::correspondence> &corrs,
                                const std::vector<bool> &inliers, size_t minimum,
                                std::vector<std::array<double, 9>> &rows);
// tail of match_features_subset (src/match/match_features.cpp:94-101) on the K1 records of one pair: ratio test in
// double, feature_match records with ORIGINAL indices, the reference's std::sort; optional cross-check flags
std::vector<opencalibration::feature_match> matches_from_top2(const std::vector<size_t> &indices_1,
                                                              const std::vector<size_t> &indices_2, const ocb_top2 *top,
                                                              const uint32_t *col, std::vector<bool> *mutual);
void rank2_from(const double *in9, double *out9);
// cv::decomposeHomographyMat(H, I, ...) restated (homography_decompose.cpp): H column-major; up to 4 solutions,
// R36 column-major 3x3 each, t12, n12; returns the number of solutions (1 for a pure rotation, else 4)
int decompose_homography_mat(const double *H9, double *R36, double *t12, double *n12);

// Eigen::Quaterniond(Matrix3d) (Shepperd's method as in Eigen/src/Geometry/Quaternion.h); R column-major
inline Eigen::Quaterniond quaternion_from_rotation(const double *R)
{
    auto m = [R](int r, int c) { return R[r + 3 * c]; };
    double q[4]; // x y z w
    double t = m(0, 0) + m(1, 1) + m(2, 2);
    if (t > 0)
    {
        t = std::sqrt(t + 1.0);
        q[3] = 0.5 * t;
        t = 0.5 / t;
        q[0] = (m(2, 1) - m(1, 2)) * t;
        q[1] = (m(0, 2) - m(2, 0)) * t;
        q[2] = (m(1, 0) - m(0, 1)) * t;
    }
    else
    {
        int i = 0;
        if (m(1, 1) > m(0, 0))
            i = 1;
        if (m(2, 2) > m(i, i))
            i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        t = std::sqrt(m(i, i) - m(j, j) - m(k, k) + 1.0);
        q[i] = 0.5 * t;
        t = 0.5 / t;
        q[3] = (m(k, j) - m(j, k)) * t;
        q[j] = (m(j, i) + m(i, j)) * t;
        q[k] = (m(k, i) + m(i, k)) * t;
    }
    return Eigen::Quaterniond(q[3], q[0], q[1], q[2]);
}

// model <-> [18] doubles of the C ABI (matrix column-major, then the inverse for the homography)
inline void pack_model(const opencalibration::homography_model &m, double *m18)
{
    for (int i = 0; i < 9; i++)
    {
        m18[i] = m.homography.data()[i];
        m18[9 + i] = m.homography_inverse.data()[i];
    }
}
inline void unpack_model(const double *m18, opencalibration::homography_model &m)
{
    for (int i = 0; i < 9; i++)
    {
        m.homography.data()[i] = m18[i];
        m.homography_inverse.data()[i] = m18[9 + i];
    }
}
inline void unpack_model(const double *m18, opencalibration::essential_matrix_model &m)
{
    for (int i = 0; i < 9; i++)
        m.essential_matrix.data()[i] = m18[i];
}
inline void unpack_model(const double *m18, opencalibration::fundamental_matrix_model &m)
{
    for (int i = 0; i < 9; i++)
        m.fundamental_matrix.data()[i] = m18[i];
}
inline void pack_model(const opencalibration::essential_matrix_model &m, double *m18)
{
    for (int i = 0; i < 9; i++)
    {
        m18[i] = m.essential_matrix.data()[i];
        m18[9 + i] = 0.0;
    }
}
inline void pack_model(const opencalibration::fundamental_matrix_model &m, double *m18)
{
    for (int i = 0; i < 9; i++)
    {
        m18[i] = m.fundamental_matrix.data()[i];
        m18[9 + i] = 0.0;
    }
}
inline int model_kind(const opencalibration::homography_model &)
{
    return OCB_MODEL_HOMOGRAPHY;
}
inline int model_kind(const opencalibration::essential_matrix_model &)
{
    return OCB_MODEL_ESSENTIAL;
}
inline int model_kind(const opencalibration::fundamental_matrix_model &)
{
    return OCB_MODEL_FUNDAMENTAL;
}
} // namespace detail
} // namespace ocb_host
