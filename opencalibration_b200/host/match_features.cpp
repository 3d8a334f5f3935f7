// match_features_subset / spatially_subsample_feature_indices with the reference's signatures
// (reference include/opencalibration/match/match_features.hpp:10-16, src/match/match_features.cpp).
// The n1 x n2 Hamming top-2 search runs on the GPU (ocb_match_top2); what stays here is what has to be the
// reference's libstdc++ behaviour bit for bit: the double-precision ratio test and the final std::sort.
#include "models_detail.hpp"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <unordered_map>

namespace
{
using opencalibration::feature_2d;
using opencalibration::feature_match;

// set[indices[k]].descriptor -> contiguous 64-byte rows (what match_features.cpp:62-66 does for set_2)
std::vector<uint64_t> pack_rows(const std::vector<feature_2d> &set, const std::vector<size_t> &indices)
{
    std::vector<uint64_t> rows(indices.size() * OCB_ROW_WORDS);
    for (size_t k = 0; k < indices.size(); k++)
        std::memcpy(&rows[k * OCB_ROW_WORDS], static_cast<const void *>(&set[indices[k]].descriptor), OCB_ROW_BYTES);
    return rows;
}

inline double as_distance(uint16_t d)
{
    // distance = count * (1.0 / DESCRIPTOR_BITS)  (match_features.cpp:79); OCB_DIST_INF stands for +infinity
    return d == OCB_DIST_INF ? std::numeric_limits<double>::infinity()
                             : d * (1.0 / feature_2d::DESCRIPTOR_BITS);
}

std::vector<feature_match> run_match(const std::vector<feature_2d> &set_1, const std::vector<feature_2d> &set_2,
                                     const std::vector<size_t> &indices_1, const std::vector<size_t> &indices_2,
                                     std::vector<bool> *mutual)
{
    const std::vector<uint64_t> q = pack_rows(set_1, indices_1), c = pack_rows(set_2, indices_2);
    std::vector<ocb_top2> top(indices_1.size());
    std::vector<uint32_t> col(mutual ? indices_2.size() : 0);
    ocb_host::detail::gpu_check(ocb_match_top2(q.data(), indices_1.size(), c.data(), indices_2.size(), top.data(),
                                               mutual ? col.data() : nullptr),
                                "ocb_match_top2");
    // Records carry the cross-check flag along; std::sort's sequence of moves depends only on the comparator's
    // answers, so sorting these by the reference's comparator gives the reference's (unstable) order.
    struct Rec
    {
        feature_match m;
        bool mutual;
    };
    std::vector<Rec> recs;
    recs.reserve(indices_1.size());
    for (size_t a = 0; a < indices_1.size(); a++)
    {
        const double best = as_distance(top[a].best_d), second = as_distance(top[a].second_d);
        if (best < 0.8 * second) // match_features.cpp:94, in double like the reference
        {
            const uint32_t k = top[a].best_k;
            recs.push_back(Rec{feature_match{indices_1[a], indices_2[k], best}, mutual && col[k] == (uint32_t)a});
        }
    }
    std::sort(recs.begin(), recs.end(),
              [](const Rec &f1, const Rec &f2) -> bool { return f1.m.distance > f2.m.distance; }); // :100-101
    std::vector<feature_match> results(recs.size());
    if (mutual)
        mutual->resize(recs.size());
    for (size_t i = 0; i < recs.size(); i++)
    {
        results[i] = recs[i].m;
        if (mutual)
            (*mutual)[i] = recs[i].mutual;
    }
    return results;
}
} // namespace

namespace opencalibration
{

std::vector<feature_match> match_features_subset(const std::vector<feature_2d> &set_1,
                                                 const std::vector<feature_2d> &set_2,
                                                 const std::vector<size_t> &indices_1,
                                                 const std::vector<size_t> &indices_2)
{
    return run_match(set_1, set_2, indices_1, indices_2, nullptr);
}

std::vector<size_t> spatially_subsample_feature_indices(const std::vector<feature_2d> &features, double spacing_pixels,
                                                        size_t count)
{
    // match_features.cpp:8-52: strongest first; keep a feature iff its nearest kept neighbour is farther than
    // `spacing_pixels` (squared-distance comparison, strict). Sequential greedy => host; the nearest-neighbour
    // query is answered from a uniform grid of kept points (cell = spacing) instead of the reference's KD-tree,
    // which yields the same minimum-distance decision.
    if (count == 0)
        count = features.size();
    if (count == 0)
        return {};
    std::vector<size_t> by_strength(count);
    for (size_t i = 0; i < count; i++)
        by_strength[i] = i;
    std::sort(by_strength.begin(), by_strength.end(),
              [&features](size_t a, size_t b) { return features[a].strength > features[b].strength; });

    std::vector<size_t> kept;
    kept.reserve(features.size() / 4);
    const double limit = spacing_pixels * spacing_pixels;
    const bool gridded = spacing_pixels > 0 && std::isfinite(spacing_pixels);
    std::unordered_map<uint64_t, std::vector<size_t>> cells;
    auto cell_key = [](int64_t cx, int64_t cy) {
        return (static_cast<uint64_t>(static_cast<uint32_t>(cx)) << 32) | static_cast<uint32_t>(cy);
    };
    auto too_close = [&](size_t idx, size_t other) {
        const double dx = features[idx].location.x() - features[other].location.x();
        const double dy = features[idx].location.y() - features[other].location.y();
        return !(dx * dx + dy * dy > limit);
    };
    for (size_t idx : by_strength)
    {
        const double x = features[idx].location.x(), y = features[idx].location.y();
        const bool finite = gridded && std::isfinite(x) && std::isfinite(y);
        bool keep = true;
        if (!kept.empty())
        {
            if (finite)
            {
                const int64_t cx = (int64_t)std::floor(x / spacing_pixels), cy = (int64_t)std::floor(y / spacing_pixels);
                for (int64_t gx = cx - 2; gx <= cx + 2 && keep; gx++)
                    for (int64_t gy = cy - 2; gy <= cy + 2 && keep; gy++)
                    {
                        auto it = cells.find(cell_key(gx, gy));
                        if (it == cells.end())
                            continue;
                        for (size_t other : it->second)
                            if (too_close(idx, other))
                            {
                                keep = false;
                                break;
                            }
                    }
            }
            else
            {
                for (size_t other : kept)
                    if (too_close(idx, other))
                    {
                        keep = false;
                        break;
                    }
            }
        }
        if (!keep)
            continue;
        if (finite)
            cells[cell_key((int64_t)std::floor(x / spacing_pixels), (int64_t)std::floor(y / spacing_pixels))].push_back(idx);
        kept.push_back(idx);
    }
    return kept;
}

} // namespace opencalibration

namespace ocb_host
{
std::vector<opencalibration::feature_match> match_features_subset_cross_checked(
    const std::vector<opencalibration::feature_2d> &set_1, const std::vector<opencalibration::feature_2d> &set_2,
    const std::vector<size_t> &indices_1, const std::vector<size_t> &indices_2, std::vector<bool> &mutual)
{
    return run_match(set_1, set_2, indices_1, indices_2, &mutual);
}
} // namespace ocb_host
