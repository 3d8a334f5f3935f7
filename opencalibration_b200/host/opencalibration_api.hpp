// The reference's C++ entry points for the matching + RANSAC path, unchanged in name, namespace, signature,
// ownership and degenerate-input behaviour, implemented on top of the C ABI of include/ocb.h (libocb.so).
//   include/opencalibration/match/match_features.hpp:10-16
//   include/opencalibration/model_inliers/ransac.hpp:15-20
// Error convention: the reference has none (no exceptions, no status codes). Here a failing GPU call throws
// std::runtime_error carrying ocb_last_error(); there is no CPU fallback.
#pragma once
#include "reference_types.hpp"

#include <ocb.h>

#include <functional>
#include <vector>

namespace opencalibration
{
std::vector<size_t> spatially_subsample_feature_indices(const std::vector<feature_2d> &features, double spacing_pixels,
                                                        size_t count = 0);

std::vector<feature_match> match_features_subset(const std::vector<feature_2d> &set_1,
                                                 const std::vector<feature_2d> &set_2,
                                                 const std::vector<size_t> &indices_1,
                                                 const std::vector<size_t> &indices_2);

template <typename Model>
double ransac(const std::vector<correspondence> &matches, Model &model, std::vector<bool> &inliers);

void assembleInliers(const std::vector<feature_match> &matches, const std::vector<bool> &inliers,
                     const std::vector<feature_2d> &source_features, const std::vector<feature_2d> &dest_features,
                     std::vector<feature_match_denormalized> &inlier_list);

// include/opencalibration/distort/distort_keypoints.hpp:17-24 (src/distort/distort_keypoints.cpp:48-103): the step
// between match and ransac in LinkStage (link_stage.cpp:87-88). Host code: ~500 matches per pair.
std::vector<correspondence> distort_keypoints(const std::vector<feature_2d> &features1,
                                              const std::vector<feature_2d> &features2,
                                              const std::vector<feature_match> &matches,
                                              const DifferentiableCameraModel<double> &model1,
                                              const DifferentiableCameraModel<double> &model2);
Eigen::Vector3d image_to_3d(const Eigen::Vector2d &keypoint, const DifferentiableCameraModel<double> &model);
} // namespace opencalibration

// Extensions that the reference does not have (kept out of its namespace).
namespace ocb_host
{
// match_features_subset + mutual-nearest-neighbour flag per returned match (cross-check): flag[m] is true iff
// the best query of match m's candidate (first minimum over indices_1 in order) is match m's query.
std::vector<opencalibration::feature_match> match_features_subset_cross_checked(
    const std::vector<opencalibration::feature_2d> &set_1, const std::vector<opencalibration::feature_2d> &set_2,
    const std::vector<size_t> &indices_1, const std::vector<size_t> &indices_2, std::vector<bool> &mutual);

// Counters of the last ransac() call on this thread (for tests: how the batched replay went).
struct RansacStats
{
    size_t iterations = 0;    // value of the reference's loop counter at exit
    size_t improvements = 0;  // times score > best_score
    size_t rejected = 0;      // SPRT rejections replayed (only improvers are replayed)
    size_t degenerate = 0;    // checkSampleDegeneracy skips among the executed iterations
    size_t scored = 0;        // hypotheses scored on the GPU (>= iterations - degenerate: batches overshoot)
    size_t gpu_calls = 0;
};
RansacStats last_ransac_stats();
// ransac<homography_model>: check, fit and score each batch of minimal samples on the device (default) or fit on the
// host and score on the device. Both give identical results; the switch exists for A/B tests and timing.
void set_ransac_device_fit(bool on);
bool ransac_device_fit();

// Many independent ransac<Model>() runs advanced in lock step (used by the batched LinkStage runner): the result of
// every job -- model, inliers, returned score -- is what ransac(*matches, *model, *inliers) gives, but each round of
// GPU work of all jobs is ONE launch instead of one per job.
template <typename Model> struct RansacJob
{
    const std::vector<opencalibration::correspondence> *matches = nullptr;
    Model *model = nullptr;
    std::vector<bool> *inliers = nullptr;
    // optional: the PROSAC ordering of src/model_inliers/ransac.cpp:83-90 (indices sorted by quality, ascending, in the
    // order the reference's std::sort leaves ties) computed elsewhere -- the batched LinkStage runner gets it from the
    // device together with the sorted match list (ocb_match_pairs_sorted); nullptr: sorted here
    const uint32_t *quality_order = nullptr;
    double result = 0; // what ransac() would have returned
    RansacStats stats;
};
// `binder` (optional) makes the jobs' correspondences resident instead of the default ocb_corr_bind_batch upload: it
// receives one ocb_corr_set per job ({rows, n, evaluation order}; n == 0 for a job that is over before it starts) and
// must leave them bound on the calling thread in that order (the batched LinkStage runner binds them straight from its
// match lists with ocb_corr_bind_batch_matches, which also computes the rays on the device).
using CorrBinder = std::function<void(const std::vector<ocb_corr_set> &)>;
template <typename Model>
void ransac_batch(std::vector<RansacJob<Model>> &jobs, int threads = 0, const CorrBinder *binder = nullptr);

// The second caller of homography_model::fitInliers / evaluate, RelaxGroup::finalize (src/relax/relax_group.cpp:
// 137-177): when the camera models changed, every edge refits its homography from its previous inliers with
//     for (int i = 0; i < 3; i++) { h.fitInliers(correspondences, inliers); h.evaluate(correspondences, inliers); }
// refit_evaluate_batch does that for all edges of a batch in lock step -- `rounds` request tables of all-inlier
// device refits + evaluations (one launch and one copy each way per round) -- and leaves in every job the model, the
// inlier vector and the last evaluate's score exactly as that loop would.
struct RefitJob
{
    const std::vector<opencalibration::correspondence> *matches = nullptr;
    opencalibration::homography_model *model = nullptr; // inlier_threshold is read, the matrices are written
    std::vector<bool> *inliers = nullptr;               // in: the previous inliers; out: the last evaluate's
    double score = 0;                                   // out: the last evaluate's return value
};
void refit_evaluate_batch(std::vector<RefitJob> &jobs, int rounds = 3);
} // namespace ocb_host
