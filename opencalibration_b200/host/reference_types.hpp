// The reference's argument / result structs for the matching + RANSAC path, as seen by the C++ mirror.
//
// Built inside the reference tree (OCB_WITH_REFERENCE_HEADERS) this header simply pulls in
// the reference's own headers, so the adapters are compiled against the real types and real Eigen:
//   include/opencalibration/types/{feature_2d,feature_match,correspondence,decomposed_pose,camera_model,
//                                  camera_relations}.hpp
//   include/opencalibration/model_inliers/{homography,essential_matrix,fundamental_matrix}_model.hpp
// Stand-alone (this repository, where Eigen is not installed) it declares layout-compatible equivalents: same
// namespace, member names, member order, sizes and alignment (static_asserts below; the expected numbers were
// measured on the reference's own translation units), with a storage-only Eigen subset.
#pragma once

#if defined(OCB_WITH_REFERENCE_HEADERS)

#include <opencalibration/model_inliers/essential_matrix_model.hpp>
#include <opencalibration/model_inliers/fundamental_matrix_model.hpp>
#include <opencalibration/model_inliers/homography_model.hpp>
#include <opencalibration/types/camera_model.hpp>
#include <opencalibration/types/camera_relations.hpp>
#include <opencalibration/types/correspondence.hpp>
#include <opencalibration/types/decomposed_pose.hpp>
#include <opencalibration/types/feature_2d.hpp>
#include <opencalibration/types/feature_match.hpp>

#else

#include <array>
#include <bitset>
#include <cmath>
#include <cstddef>
#include <vector>

namespace Eigen
{
// Storage-only subset: fixed-size, column-major, same size/alignment as Eigen 3.4's dense types.
template <int N> struct OcbVec
{
    double v[N];
    double &operator[](std::size_t i) { return v[i]; }
    const double &operator[](std::size_t i) const { return v[i]; }
    double &operator()(std::size_t i) { return v[i]; }
    const double &operator()(std::size_t i) const { return v[i]; }
    double &x() { return v[0]; }
    double &y() { return v[1]; }
    double &z() { return v[2]; }
    const double &x() const { return v[0]; }
    const double &y() const { return v[1]; }
    const double &z() const { return v[2]; }
    const double *data() const { return v; }
    double *data() { return v; }
    bool operator==(const OcbVec &o) const
    {
        bool same = true;
        for (int i = 0; i < N; i++)
            same = same && (v[i] == o.v[i]);
        return same;
    }
    bool allNaN() const
    {
        bool a = true;
        for (int i = 0; i < N; i++)
            a = a && std::isnan(v[i]);
        return a;
    }
};
struct alignas(16) Vector2d : OcbVec<2>
{
    Vector2d() : OcbVec<2>{{NAN, NAN}} {}
    Vector2d(double a, double b) : OcbVec<2>{{a, b}} {}
};
struct Vector3d : OcbVec<3>
{
    Vector3d() : OcbVec<3>{{NAN, NAN, NAN}} {}
    Vector3d(double a, double b, double c) : OcbVec<3>{{a, b, c}} {}
};
struct alignas(16) Vector4d : OcbVec<4>
{
    Vector4d() : OcbVec<4>{{NAN, NAN, NAN, NAN}} {}
    Vector4d(double a, double b, double c, double d) : OcbVec<4>{{a, b, c, d}} {}
};
struct Matrix3d
{
    double m[9]; // column-major: (r,c) -> m[r + 3*c]
    Matrix3d()
    {
        for (double &e : m)
            e = NAN;
    }
    static Matrix3d Constant(double value)
    {
        Matrix3d r;
        for (double &e : r.m)
            e = value;
        return r;
    }
    static Matrix3d Identity()
    {
        Matrix3d r = Constant(0.0);
        r.m[0] = r.m[4] = r.m[8] = 1.0;
        return r;
    }
    double &operator()(std::size_t r, std::size_t c) { return m[r + 3 * c]; }
    const double &operator()(std::size_t r, std::size_t c) const { return m[r + 3 * c]; }
    const double *data() const { return m; }
    double *data() { return m; }
    bool operator==(const Matrix3d &o) const
    {
        bool same = true;
        for (int i = 0; i < 9; i++)
            same = same && (m[i] == o.m[i]);
        return same;
    }
};
struct Quaterniond
{
    Vector4d xyzw; // Eigen's coefficient order
    Quaterniond() {}
    Quaterniond(double w, double x, double y, double z) : xyzw(x, y, z, w) {}
    const Vector4d &coeffs() const { return xyzw; }
    Vector4d &coeffs() { return xyzw; }
    double w() const { return xyzw[3]; }
    double x() const { return xyzw[0]; }
    double y() const { return xyzw[1]; }
    double z() const { return xyzw[2]; }
};
} // namespace Eigen

namespace opencalibration
{
// include/opencalibration/types/feature_2d.hpp:9-21
struct feature_2d
{
    static constexpr int DESCRIPTOR_BITS = 486;
    Eigen::Vector2d location;
    float strength = 0;
    std::bitset<DESCRIPTOR_BITS> descriptor;
};
// include/opencalibration/types/feature_match.hpp:10-21
struct feature_match
{
    size_t feature_index_1;
    size_t feature_index_2;
    double distance;
    bool operator==(const feature_match &o) const
    {
        return feature_index_1 == o.feature_index_1 && feature_index_2 == o.feature_index_2 && distance == o.distance;
    }
};
// include/opencalibration/types/feature_match.hpp:24-36
struct feature_match_denormalized
{
    Eigen::Vector2d pixel_1, pixel_2;
    size_t feature_index_1, feature_index_2, match_index;
};
// include/opencalibration/types/correspondence.hpp:8-13
struct correspondence
{
    Eigen::Vector3d measurement1;
    Eigen::Vector3d measurement2;
    double quality{0};
};
// include/opencalibration/types/decomposed_pose.hpp:7-20
struct decomposed_pose
{
    Eigen::Quaterniond orientation{NAN, NAN, NAN, NAN};
    Eigen::Vector3d position{NAN, NAN, NAN};
    int score{0};
};

// include/opencalibration/types/camera_relations.hpp:13-35 -- the edge payload LinkStage fills
struct camera_relations
{
    std::vector<feature_match_denormalized> inlier_matches;
    std::vector<feature_match> matches;
    Eigen::Matrix3d ransac_relation = Eigen::Matrix3d::Constant(NAN);
    enum class RelationType
    {
        HOMOGRAPHY,
        FUNDAMENTAL_MATRIX,
        UNKNOWN
    } relationType = RelationType::UNKNOWN;
    std::array<decomposed_pose, 4> relative_poses;
};
// include/opencalibration/types/camera_model.hpp:10-60 (the members image_to_3d reads; T = double only here)
enum class ProjectionType
{
    PLANAR,
    UNKNOWN
};
enum class CameraModelTag
{
    FORWARD,
    INVERSE
};
template <typename T, CameraModelTag tag> struct DifferentiableCameraModelBase
{
    size_t pixels_rows = 0;
    size_t pixels_cols = 0;
    T focal_length_pixels = T(0);
    Eigen::Vector2d principle_point{0, 0};
    Eigen::Vector3d radial_distortion{0, 0, 0};
    Eigen::Vector2d tangential_distortion{0, 0};
    ProjectionType projection_type = ProjectionType::PLANAR;
};
template <typename T> using DifferentiableCameraModel = DifferentiableCameraModelBase<T, CameraModelTag::FORWARD>;
template <typename T>
using InverseDifferentiableCameraModel = DifferentiableCameraModelBase<T, CameraModelTag::INVERSE>;

// include/opencalibration/model_inliers/homography_model.hpp:14-34
struct homography_model
{
    homography_model();
    static constexpr size_t MINIMUM_POINTS = 4;
    void fit(const std::vector<correspondence> &corrs, const std::array<size_t, MINIMUM_POINTS> &initial_indices);
    void fitInliers(const std::vector<correspondence> &corrs, const std::vector<bool> &inliers);
    double evaluate(const std::vector<correspondence> &corrs, std::vector<bool> &inliers);
    bool decompose(const std::vector<correspondence> &corrs, const std::vector<bool> &inliers,
                   std::array<decomposed_pose, 4> &poses);
    double error(const correspondence &cor);
    static bool checkSampleDegeneracy(const std::vector<correspondence> &corrs,
                                      const std::array<size_t, MINIMUM_POINTS> &indices);
    double inlier_threshold = 0.005;
    Eigen::Matrix3d homography;
    Eigen::Matrix3d homography_inverse;
};
// include/opencalibration/model_inliers/essential_matrix_model.hpp:15-33
struct essential_matrix_model
{
    essential_matrix_model();
    static constexpr size_t MINIMUM_POINTS = 5;
    void fit(const std::vector<correspondence> &corrs, const std::array<size_t, MINIMUM_POINTS> &initial_indices);
    void fitInliers(const std::vector<correspondence> &corrs, const std::vector<bool> &inliers);
    double evaluate(const std::vector<correspondence> &corrs, std::vector<bool> &inliers);
    double error(const correspondence &cor);
    bool decompose(const std::vector<correspondence> &corrs, const std::vector<bool> &inliers,
                   std::array<decomposed_pose, 4> &poses);
    double inlier_threshold{0.01};
    Eigen::Matrix3d essential_matrix;
};
// include/opencalibration/model_inliers/fundamental_matrix_model.hpp:15-31
struct fundamental_matrix_model
{
    fundamental_matrix_model();
    static constexpr size_t MINIMUM_POINTS = 8;
    void fit(const std::vector<correspondence> &corrs, const std::array<size_t, MINIMUM_POINTS> &initial_indices);
    void fitInliers(const std::vector<correspondence> &corrs, const std::vector<bool> &inliers);
    double evaluate(const std::vector<correspondence> &corrs, std::vector<bool> &inliers);
    double error(const correspondence &cor);
    void checkDegeneracy(const std::vector<correspondence> &corrs, std::vector<bool> &inliers);
    double inlier_threshold = 0.01;
    Eigen::Matrix3d fundamental_matrix;
};
} // namespace opencalibration

#endif // OCB_WITH_REFERENCE_HEADERS

namespace opencalibration
{
static_assert(sizeof(std::bitset<feature_2d::DESCRIPTOR_BITS>) == 64, "descriptor row = 64 bytes");
static_assert(sizeof(feature_2d) == 96 && offsetof(feature_2d, descriptor) == 24, "feature_2d layout");
static_assert(sizeof(feature_match) == 24, "feature_match layout");
static_assert(sizeof(correspondence) == 56, "correspondence = 7 doubles");
} // namespace opencalibration
