// TEST INFRASTRUCTURE ONLY -- the model member functions that src/model_inliers/ransac.cpp (compiled in
// place from /root/reference into oracle/_ref/) leaves undefined. The reference's own definitions
// (src/model_inliers/{homography,essential_matrix,fundamental_matrix}_model.cpp) call Eigen algorithms
// (fullPivLu, jacobiSvd, inverse, hnormalized) and OpenCV, neither of which exists in this image, so they
// delegate to the restatement in oc_oracle.cpp. Compiled against the reference's own headers (with the
// storage-only Eigen stand-in), so signatures and struct layouts are the reference's.
#include <opencalibration/model_inliers/ransac.hpp>

#include "oc_oracle.hpp"

#include <cmath>
#include <cstring>

using namespace opencalibration;

static_assert(sizeof(correspondence) == sizeof(oc_oracle::Corr), "correspondence layout");
static_assert(sizeof(correspondence) == 56, "correspondence is 7 doubles");

namespace
{
const oc_oracle::Corr *as_corr(const std::vector<correspondence> &v)
{
    return reinterpret_cast<const oc_oracle::Corr *>(v.data());
}
oc_oracle::Model to_model(const homography_model &h)
{
    oc_oracle::Model m(oc_oracle::MODEL_HOMOGRAPHY);
    m.thr = h.inlier_threshold;
    std::memcpy(m.M, h.homography.d, sizeof m.M);
    std::memcpy(m.Minv, h.homography_inverse.d, sizeof m.Minv);
    return m;
}
void from_model(const oc_oracle::Model &m, homography_model &h)
{
    std::memcpy(h.homography.d, m.M, sizeof m.M);
    std::memcpy(h.homography_inverse.d, m.Minv, sizeof m.Minv);
}
template <typename T> oc_oracle::Model to_model_epi(const T &e, const Eigen::Matrix3d &M, int kind)
{
    oc_oracle::Model m(kind);
    m.thr = e.inlier_threshold;
    std::memcpy(m.M, M.d, sizeof m.M);
    return m;
}
} // namespace

namespace opencalibration
{
// ---- homography_model ---------------------------------------------------------------------------
homography_model::homography_model()
    : homography(Eigen::Matrix3d::Constant(NAN)), homography_inverse(Eigen::Matrix3d::Constant(NAN))
{
}
void homography_model::fit(const std::vector<correspondence> &corrs,
                           const std::array<size_t, MINIMUM_POINTS> &initial_indices)
{
    oc_oracle::Model m = to_model(*this);
    oc_oracle::fit(m, as_corr(corrs), initial_indices.data());
    from_model(m, *this);
}
void homography_model::fitInliers(const std::vector<correspondence> &corrs, const std::vector<bool> &inliers)
{
    oc_oracle::Model m = to_model(*this);
    oc_oracle::fit_inliers(m, as_corr(corrs), corrs.size(), inliers);
    from_model(m, *this);
}
double homography_model::evaluate(const std::vector<correspondence> &corrs, std::vector<bool> &inliers)
{
    return oc_oracle::evaluate(to_model(*this), as_corr(corrs), corrs.size(), inliers);
}
double homography_model::error(const correspondence &cor)
{
    return oc_oracle::error(to_model(*this), reinterpret_cast<const oc_oracle::Corr &>(cor));
}
bool homography_model::checkSampleDegeneracy(const std::vector<correspondence> &corrs,
                                             const std::array<size_t, MINIMUM_POINTS> &indices)
{
    return oc_oracle::check_sample_degeneracy_h(as_corr(corrs), indices.data());
}

// ---- essential_matrix_model ---------------------------------------------------------------------
essential_matrix_model::essential_matrix_model() : essential_matrix(Eigen::Matrix3d::Constant(NAN))
{
}
void essential_matrix_model::fit(const std::vector<correspondence> &corrs,
                                 const std::array<size_t, MINIMUM_POINTS> &initial_indices)
{
    oc_oracle::Model m = to_model_epi(*this, essential_matrix, oc_oracle::MODEL_ESSENTIAL);
    oc_oracle::fit(m, as_corr(corrs), initial_indices.data());
    std::memcpy(essential_matrix.d, m.M, sizeof m.M);
}
void essential_matrix_model::fitInliers(const std::vector<correspondence> &corrs, const std::vector<bool> &inliers)
{
    oc_oracle::Model m = to_model_epi(*this, essential_matrix, oc_oracle::MODEL_ESSENTIAL);
    oc_oracle::fit_inliers(m, as_corr(corrs), corrs.size(), inliers);
    std::memcpy(essential_matrix.d, m.M, sizeof m.M);
}
double essential_matrix_model::evaluate(const std::vector<correspondence> &corrs, std::vector<bool> &inliers)
{
    return oc_oracle::evaluate(to_model_epi(*this, essential_matrix, oc_oracle::MODEL_ESSENTIAL), as_corr(corrs),
                               corrs.size(), inliers);
}
double essential_matrix_model::error(const correspondence &cor)
{
    return oc_oracle::error(to_model_epi(*this, essential_matrix, oc_oracle::MODEL_ESSENTIAL),
                            reinterpret_cast<const oc_oracle::Corr &>(cor));
}

// ---- fundamental_matrix_model -------------------------------------------------------------------
fundamental_matrix_model::fundamental_matrix_model() : fundamental_matrix(Eigen::Matrix3d::Constant(NAN))
{
}
void fundamental_matrix_model::fit(const std::vector<correspondence> &corrs,
                                   const std::array<size_t, MINIMUM_POINTS> &initial_indices)
{
    oc_oracle::Model m = to_model_epi(*this, fundamental_matrix, oc_oracle::MODEL_FUNDAMENTAL);
    oc_oracle::fit(m, as_corr(corrs), initial_indices.data());
    std::memcpy(fundamental_matrix.d, m.M, sizeof m.M);
}
void fundamental_matrix_model::fitInliers(const std::vector<correspondence> &corrs, const std::vector<bool> &inliers)
{
    oc_oracle::Model m = to_model_epi(*this, fundamental_matrix, oc_oracle::MODEL_FUNDAMENTAL);
    oc_oracle::fit_inliers(m, as_corr(corrs), corrs.size(), inliers);
    std::memcpy(fundamental_matrix.d, m.M, sizeof m.M);
}
double fundamental_matrix_model::evaluate(const std::vector<correspondence> &corrs, std::vector<bool> &inliers)
{
    return oc_oracle::evaluate(to_model_epi(*this, fundamental_matrix, oc_oracle::MODEL_FUNDAMENTAL), as_corr(corrs),
                               corrs.size(), inliers);
}
double fundamental_matrix_model::error(const correspondence &cor)
{
    return oc_oracle::error(to_model_epi(*this, fundamental_matrix, oc_oracle::MODEL_FUNDAMENTAL),
                            reinterpret_cast<const oc_oracle::Corr &>(cor));
}
void fundamental_matrix_model::checkDegeneracy(const std::vector<correspondence> &corrs, std::vector<bool> &inliers)
{
    oc_oracle::Model m = to_model_epi(*this, fundamental_matrix, oc_oracle::MODEL_FUNDAMENTAL);
    oc_oracle::check_degeneracy_f(m, as_corr(corrs), corrs.size(), inliers);
    std::memcpy(fundamental_matrix.d, m.M, sizeof m.M);
}
} // namespace opencalibration
