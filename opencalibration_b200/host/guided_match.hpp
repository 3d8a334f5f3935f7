// Guided matcher of the dense stage (reference src/dense/dense_stereo.cpp:244-281), batched: see guided_match.cpp.
#pragma once
#include "reference_types.hpp"

#include <cstddef>
#include <vector>

namespace ocb_host
{
// The (source feature, candidate image) visits of densifyMesh's inner loops for ONE candidate image, in CSR form:
// visit l looks at source feature query_feature[l] and at the candidate-image features
// nearby[begin[l] .. begin[l+1]) -- the payloads of ft_searcher.search(predicted, 150^2, max) in result order
// (dense_stereo.cpp:244-246; jk-tree returns them by ascending distance).
struct GuidedLists
{
    std::vector<size_t> query_feature; // [n_lists] index into the source image's feature vector
    std::vector<size_t> begin;         // [n_lists + 1]
    std::vector<size_t> nearby;        // candidate-image feature indices
};
// One accepted visit (the reference pushes {src_id, measurementId(cand_nid, best_feat_idx)}, :277-280).
struct GuidedMatch
{
    size_t list;
    size_t query_feature;
    size_t candidate_feature; // best_feat_idx
    double best_distance;     // best_dist
    double second_distance;   // second_best_dist (+inf for a single candidate)
};
// Top-2 scan (:251-273) on the GPU + acceptance rule (:275-276), accepted visits in list order.
// Throws std::runtime_error when the GPU call fails (no CPU fallback).
std::vector<GuidedMatch> match_features_guided(const std::vector<opencalibration::feature_2d> &source,
                                               const std::vector<opencalibration::feature_2d> &candidates,
                                               const GuidedLists &lists);
} // namespace ocb_host
