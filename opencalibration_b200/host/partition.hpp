// Overlap-graph partition of a LinkStage pair list over several GPUs.
//
// Directed image pairs are independent units of work (reference src/pipeline/link_stage.cpp:75-112). Images are
// ordered along a Hilbert curve over their positions -- the same curve as the reference's xy2d helper
// (include/opencalibration/types/hilbert.hpp:8-27) -- and the curve is cut into `world` contiguous runs balanced by
// the number of pairs each image sources; a pair belongs to the part that owns its SOURCE image, and a part keeps
// resident its own images plus the "halo" images its pairs reference. There is no data-path collective: the only
// cross-part step is putting the per-pair results back into the serial pair order, which is what
// LinkStage::finalize does with its runners' results (link_stage.cpp:119-131).
#pragma once
#include "link_batch.hpp"

#include <cstddef>
#include <cstdint>
#include <vector>

namespace ocb_host
{
// Position of integer cell (x, y) along the Hilbert curve over an order x order grid (order = power of two).
uint32_t hilbert_index(int order, int x, int y);
// Permutation of 0 .. n-1 along the curve: positions are scaled to a 1024 x 1024 grid over their bounding box; ties
// (same cell) keep index order.
std::vector<size_t> hilbert_order(const double *xy, size_t n);

struct PairShard
{
    std::vector<size_t> owned_images; // ascending: images whose outgoing pairs this part matches
    std::vector<size_t> halo_images;  // ascending: other images those pairs reference (resident too)
    std::vector<size_t> pair_ids;     // ascending indices into the pair list (= serial order)
};
// xy: [n_images][2] positions (nullptr: the images' index order stands in for the curve).
std::vector<PairShard> partition_pairs(const double *xy, size_t n_images, const std::vector<LinkPair> &pairs,
                                       size_t world);

// link_pairs over `n_devices` GPUs of this process (devices 0 .. n_devices-1): the pair list is partitioned as above,
// one submission thread per device runs the single-device runner on its part (its own and its halo images resident
// on that device), the per-pair host tail of all parts shares the process's cores, and the results come back in pair
// order -- identical to link_pairs(images, pairs). This is the entry point a single-process caller such as
// pipeline_runner's LinkStage (app/pipeline_runner.cpp:313-326, src/pipeline/pipeline.cpp:543-560) uses on a
// multi-GPU box. xy as in partition_pairs.
std::vector<opencalibration::camera_relations> link_pairs_multi(const std::vector<LinkImage> &images,
                                                                const std::vector<LinkPair> &pairs, const double *xy,
                                                                int n_devices, const LinkOptions &options = LinkOptions(),
                                                                LinkStats *stats = nullptr);
} // namespace ocb_host
