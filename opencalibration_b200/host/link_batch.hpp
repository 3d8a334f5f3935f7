// Batched LinkStage runner: what the closures of LinkStage::get_runners do for every
// (image, neighbour) pair (reference src/pipeline/link_stage.cpp:63-65,75-112), for a whole batch of pairs at once.
//   reference: one closure per pair on an OpenMP worker: subsample -> match_features_subset -> distort_keypoints ->
//              ransac<homography_model> -> decompose -> assembleInliers, all on that worker's core;
//   here:      the packed descriptor rows of every image are uploaded ONCE (ocb_register_descriptors), in the order
//              of first use and ahead of the matching; the pairs are matched in large submissions (ocb_match_pairs:
//              one K1 launch per submission, followed on the device by the ratio test and an order-preserving
//              compaction, K5), and the per-pair tail (the reference's std::sort, rays on the device (K6), RANSAC
//              with device fits, refits and scoring, decomposition, inlier assembly) runs on OpenMP workers while
//              the next submission is on the GPU.
// Results are returned in pair order (the order LinkStage::finalize restores, link_stage.cpp:119-131) and are
// identical to running the reference-signature functions of opencalibration_api.hpp pair by pair.
#pragma once
#include "opencalibration_api.hpp"

#include <cstddef>
#include <vector>

namespace ocb_host
{
// What a LinkStage closure reads from an `image` node (include/opencalibration/types/image.hpp:18-48).
struct LinkImage
{
    const std::vector<opencalibration::feature_2d> *features = nullptr;
    size_t num_sparse_features = 0; // link_stage.cpp:63-65: subsample only the sparse prefix (0 = all)
    opencalibration::DifferentiableCameraModel<double> model;
};
struct LinkPair
{
    size_t image_1; // node_id: the query side
    size_t image_2; // match_node_id: the candidate side
};
struct LinkOptions
{
    int threads = 0;                    // OpenMP workers of the per-pair tail (0 = all cores)
    size_t pairs_per_submission = 256;  // pairs per ocb_match_pairs call
    double coarse_spacing_pixels = 40.0; // link_stage.cpp:62
    bool run_ransac = true;             // false: stop after the match lists (relations.matches only)
    int tail_workers = 4;               // chunks whose tails (incl. the lock-step RANSAC rounds) run concurrently
    size_t first_submission = 32;       // pairs of the first submission; later ones double up to pairs_per_submission
    // true: ratio test + compaction (K5) and the pixel -> ray step (K6) run on the device, so that only the surviving
    // matches cross PCIe and the host keeps the reference's std::sort, the RANSAC control flow, decompose and
    // assembleInliers. false: the K1 records of every query come back and the host does the ratio test and the rays
    // (the round-1 path, kept for A/B tests: both give identical relations).
    bool device_tail = true;
    // With the device tail: also run the reference's two std::sort calls on the device (K7, libstdc++'s introsort
    // replayed step for step: match list by distance, PROSAC pool by quality) instead of on the tail workers. Exact
    // either way (tests/test_sort_replay.py, tests/test_gpu_link_tail.py). Off by default: the replay is sequential per
    // pair (~3 ms per submission on one warp per pair) and measured SLOWER end to end than sorting on the host, on one
    // GPU (-3 .. -9 % pairs/s) and on eight.
    bool device_sort = false;
    // Optional flat copy of every pair's final match list, written by the tail workers as the submissions finish (what
    // a rank contributes to the host gather of the match lists, link_stage.cpp:119-131): 12-byte records
    // {feature_index_1, feature_index_2, integer Hamming distance} (distance = d * (1.0 / 486) exactly) into
    // packed_out, which holds packed_capacity records (e.g. a region of shared memory); pair p's records are
    // packed_out[3 * packed_offsets[p] ..) for packed_counts[p] records. The lists of one submission are contiguous; the
    // submissions land in the order in which their tails finish. All three arrays are the caller's (offsets / counts:
    // one entry per pair).
    uint32_t *packed_out = nullptr;
    size_t packed_capacity = 0;
    uint64_t *packed_offsets = nullptr;
    uint64_t *packed_counts = nullptr;
};
struct LinkStats
{
    double seconds_subsample_upload = 0, seconds_setup = 0, seconds_match_gpu = 0, seconds_tail = 0,
           seconds_release = 0, seconds_total = 0;
    size_t comparisons = 0, matches = 0, ransac_inliers = 0;
};
// relations[p] is what the closure of pair p would have stored in its edge_payload (link_stage.cpp:95-111).
// With run_ransac == false only relations[p].matches is filled.
std::vector<opencalibration::camera_relations> link_pairs(const std::vector<LinkImage> &images,
                                                          const std::vector<LinkPair> &pairs,
                                                          const LinkOptions &options = LinkOptions(),
                                                          LinkStats *stats = nullptr);
} // namespace ocb_host
