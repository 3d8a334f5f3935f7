// distort_keypoints / image_to_3d with the reference's signatures (reference
// include/opencalibration/distort/distort_keypoints.hpp:17-24, src/distort/distort_keypoints.cpp:48-103):
// pixel -> unit ray through the camera model, the only step between match_features_subset and ransac in
// LinkStage (src/pipeline/link_stage.cpp:87-88). It stays on the host: a pair yields ~500 matches, and the match
// list has to pass through the host anyway for the reference's double-precision ratio test and std::sort.
//
//   * zero distortion (the LinkStage case: models come from EXIF with no distortion until relax refines them):
//     ray = ((kp - pp) / f).homogeneous().normalized(), each operation a single IEEE double operation in the order
//     of the Eigen expression (:67,96-97): squaredNorm = (x*x + y*y) + 1*1, every component divided by its sqrt;
//   * non-zero distortion (:74-91): the reference inverts distortProjectedRay (distort_keypoints.hpp:27-43) with
//     ceres::TinySolver, a dense Levenberg-Marquardt on 2 parameters, from the unprojected point, with
//     parameter_tolerance 1e-2 / (|pp| + f), gradient_tolerance 1e-2 of that, <= 10 iterations, cost_threshold 1e-16.
//     Ceres is an external, un-vendored dependency (find_package(Ceres), CMakeLists.txt:38); the published
//     algorithm (Jacobi-scaled normal equations, u *= max(1/3, 1 - (2 rho - 1)^3) on success, u *= v, v *= 2 on
//     failure, initial radius 1e4) is restated with the analytic Jacobian in place of autodiff. Pinned like the
//     reference pins it: round trip through the forward model to 1e-2 px (test/test_distort.cpp:45-67).
#include "models_detail.hpp"

#include <algorithm>
#include <cmath>

namespace
{
struct Distortion
{
    double k[3]; // radial
    double p[2]; // tangential
};

// distortProjectedRay (distort_keypoints.hpp:27-43) and, optionally, its 2x2 Jacobian (row-major)
void distort_ray(const Distortion &d, double x, double y, double *out, double *J)
{
    const double r2 = x * x + y * y, r4 = r2 * r2, r6 = r4 * r2;
    const double radial = 1.0 + (d.k[0] * r2 + d.k[1] * r4 + d.k[2] * r6);
    const double xy = x * y;
    out[0] = radial * x + 2.0 * xy * d.p[0] + d.p[1] * (r2 + 2.0 * x * x);
    out[1] = radial * y + 2.0 * xy * d.p[1] + d.p[0] * (r2 + 2.0 * y * y);
    if (J)
    {
        const double dr = d.k[0] + 2.0 * d.k[1] * r2 + 3.0 * d.k[2] * r4; // d radial / d r2
        const double drx = dr * 2.0 * x, dry = dr * 2.0 * y;
        J[0] = radial + x * drx + 2.0 * y * d.p[0] + d.p[1] * (2.0 * x + 4.0 * x);
        J[1] = x * dry + 2.0 * x * d.p[0] + d.p[1] * (2.0 * y);
        J[2] = y * drx + 2.0 * y * d.p[1] + d.p[0] * (2.0 * x);
        J[3] = radial + y * dry + 2.0 * x * d.p[1] + d.p[0] * (2.0 * y + 4.0 * y);
    }
}

// ceres::TinySolver<..., 2, 2>::Solve restated for residual(x) = target - distort(x)
void undistort_lm(const Distortion &d, const double *target, double parameter_tolerance, double *x)
{
    const double gradient_tolerance = parameter_tolerance * 1e-2, cost_threshold = 1e-16;
    const int max_num_iterations = 10;
    double f[2], J[4], scale[2], jtj[4], g[2], cost = 0;
    auto update = [&]() { // residuals, scaled Jacobian, normal equations at x; returns max |gradient|
        double dist[2], Jd[4];
        distort_ray(d, x[0], x[1], dist, Jd);
        f[0] = target[0] - dist[0], f[1] = target[1] - dist[1];
        for (int i = 0; i < 4; i++)
            J[i] = -Jd[i];
        for (int c = 0; c < 2; c++)
        {
            scale[c] = 1.0 / (1.0 + std::sqrt(J[c] * J[c] + J[2 + c] * J[2 + c]));
            J[c] *= scale[c], J[2 + c] *= scale[c];
        }
        jtj[0] = J[0] * J[0] + J[2] * J[2], jtj[1] = jtj[2] = J[0] * J[1] + J[2] * J[3];
        jtj[3] = J[1] * J[1] + J[3] * J[3];
        g[0] = -(J[0] * f[0] + J[2] * f[1]), g[1] = -(J[1] * f[0] + J[3] * f[1]); // g = J^T * (-f)
        cost = 0.5 * (f[0] * f[0] + f[1] * f[1]);
        return std::max(std::abs(g[0]), std::abs(g[1]));
    };
    if (update() < gradient_tolerance || cost < cost_threshold)
        return;
    double u = 1.0 / 1e4, v = 2.0;
    for (int it = 1; it < max_num_iterations; it++)
    {
        const double a = jtj[0] + u * std::min(std::max(jtj[0], 1e-6), 1e32), b = jtj[1];
        const double c = jtj[3] + u * std::min(std::max(jtj[3], 1e-6), 1e32);
        const double det = a * c - b * b;
        const double s0 = (c * g[0] - b * g[1]) / det, s1 = (a * g[1] - b * g[0]) / det; // lm_step
        const double dx0 = scale[0] * s0, dx1 = scale[1] * s1;
        const double xnorm = std::sqrt(x[0] * x[0] + x[1] * x[1]);
        if (std::sqrt(dx0 * dx0 + dx1 * dx1) < parameter_tolerance * (xnorm + parameter_tolerance))
            return;
        const double xn[2] = {x[0] + dx0, x[1] + dx1};
        double dist[2];
        distort_ray(d, xn[0], xn[1], dist, nullptr);
        const double fn0 = target[0] - dist[0], fn1 = target[1] - dist[1];
        const double cost_change = 2.0 * cost - (fn0 * fn0 + fn1 * fn1);
        const double model_change = s0 * (2.0 * g[0] - (jtj[0] * s0 + jtj[1] * s1)) +
                                    s1 * (2.0 * g[1] - (jtj[2] * s0 + jtj[3] * s1));
        const double rho = cost_change / model_change;
        if (rho > 0)
        {
            x[0] = xn[0], x[1] = xn[1];
            if (update() < gradient_tolerance || cost < cost_threshold)
                return;
            const double tmp = 2.0 * rho - 1.0;
            u = u * std::max(1.0 / 3.0, 1.0 - tmp * tmp * tmp);
            v = 2.0;
            continue;
        }
        u *= v;
        v *= 2.0;
    }
}
} // namespace

namespace opencalibration
{
Eigen::Vector3d image_to_3d(const Eigen::Vector2d &keypoint, const DifferentiableCameraModel<double> &model)
{
    const double f = model.focal_length_pixels;
    const double unprojected[2] = {(keypoint[0] - model.principle_point[0]) / f,
                                   (keypoint[1] - model.principle_point[1]) / f};
    double und[2] = {unprojected[0], unprojected[1]};
    const Distortion d{{model.radial_distortion[0], model.radial_distortion[1], model.radial_distortion[2]},
                       {model.tangential_distortion[0], model.tangential_distortion[1]}};
    if (d.k[0] != 0 || d.k[1] != 0 || d.k[2] != 0 || d.p[0] != 0 || d.p[1] != 0)
    {
        const double pp_norm = std::sqrt(model.principle_point[0] * model.principle_point[0] +
                                         model.principle_point[1] * model.principle_point[1]);
        undistort_lm(d, unprojected, 1e-2 / (pp_norm + f), und);
    }
    Eigen::Vector3d ray; // left unset for ProjectionType::UNKNOWN, like the reference (:93-101)
    if (model.projection_type == ProjectionType::PLANAR)
    {
        const double z = (und[0] * und[0] + und[1] * und[1]) + 1.0 * 1.0;
        if (z > 0) // Eigen's normalized()
        {
            const double n = std::sqrt(z);
            ray = Eigen::Vector3d(und[0] / n, und[1] / n, 1.0 / n);
        }
        else
            ray = Eigen::Vector3d(und[0], und[1], 1.0);
    }
    return ray;
}

std::vector<correspondence> distort_keypoints(const std::vector<feature_2d> &features1,
                                              const std::vector<feature_2d> &features2,
                                              const std::vector<feature_match> &matches,
                                              const DifferentiableCameraModel<double> &model1,
                                              const DifferentiableCameraModel<double> &model2)
{
    std::vector<correspondence> distorted;
    distorted.reserve(matches.size());
    for (const feature_match &m : matches)
    {
        correspondence cor;
        cor.measurement1 = image_to_3d(features1[m.feature_index_1].location, model1);
        cor.measurement2 = image_to_3d(features2[m.feature_index_2].location, model2);
        cor.quality = m.distance;
        distorted.push_back(cor);
    }
    return distorted;
}
} // namespace opencalibration
