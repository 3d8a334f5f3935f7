// match_features_subset / spatially_subsample_feature_indices with the reference's signatures
// (reference include/opencalibration/match/match_features.hpp:10-16, src/match/match_features.cpp).
// The n1 x n2 Hamming top-2 search runs on the GPU (ocb_match_top2); what stays here is what has to be the
// reference's libstdc++ behaviour bit for bit: the double-precision ratio test and the final std::sort.
#include "models_detail.hpp"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>

namespace
{
using opencalibration::feature_2d;
using opencalibration::feature_match;

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
    // set[indices[k]].descriptor is gathered by the library straight into its page-locked staging area
    // (what match_features.cpp:62-66 does for set_2 into `packed_2`)
    static_assert(sizeof(feature_2d::descriptor) == OCB_ROW_BYTES, "bitset<486> is 8 x 64-bit words");
    const size_t n1 = indices_1.size(), n2 = indices_2.size();
    std::vector<ocb_top2> top(n1);
    std::vector<uint32_t> col(mutual ? n2 : 0);
    const void *rows1 = set_1.empty() ? nullptr : static_cast<const void *>(&set_1[0].descriptor);
    const void *rows2 = set_2.empty() ? nullptr : static_cast<const void *>(&set_2[0].descriptor);
    ocb_host::detail::gpu_check(ocb_match_top2_strided(rows1, sizeof(feature_2d), indices_1.data(), n1, rows2,
                                                       sizeof(feature_2d), indices_2.data(), n2, top.data(),
                                                       mutual ? col.data() : nullptr),
                                "ocb_match_top2_strided");
    return ocb_host::detail::matches_from_top2(indices_1, indices_2, top.data(), mutual ? col.data() : nullptr, mutual);
}
} // namespace

namespace ocb_host
{
namespace detail
{
std::vector<feature_match> matches_from_top2(const std::vector<size_t> &indices_1, const std::vector<size_t> &indices_2,
                                             const ocb_top2 *top, const uint32_t *col, std::vector<bool> *mutual)
{
    // Ratio test in double like the reference (:94), then the reference's std::sort (:100-101). The sort runs on
    // compact (query position, integer distance) records: std::sort's sequence of comparisons and moves depends
    // only on the comparator's answers, and a.d > b.d <=> a.d * (1.0 / 486) > b.d * (1.0 / 486) for these
    // integers, so the permutation is the one the reference's sort produces on its 24-byte records.
    const size_t n1 = indices_1.size();
    struct Rec
    {
        uint32_t a;
        uint32_t d;
    };
    std::vector<Rec> recs;
    recs.reserve(n1);
    for (size_t a = 0; a < n1; a++)
    {
        const double best = as_distance(top[a].best_d), second = as_distance(top[a].second_d);
        if (best < 0.8 * second)
            recs.push_back(Rec{(uint32_t)a, top[a].best_d});
    }
    std::sort(recs.begin(), recs.end(), [](const Rec &f1, const Rec &f2) -> bool { return f1.d > f2.d; });
    std::vector<feature_match> results(recs.size());
    if (mutual)
        mutual->resize(recs.size());
    for (size_t i = 0; i < recs.size(); i++)
    {
        const uint32_t a = recs[i].a, k = top[a].best_k;
        results[i] = feature_match{indices_1[a], indices_2[k], as_distance(top[a].best_d)};
        if (mutual)
            (*mutual)[i] = col && col[k] == a;
    }
    return results;
}
} // namespace detail
} // namespace ocb_host

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
    // query is answered from a uniform grid of kept points instead of the reference's KD-tree, which yields the
    // same minimum-distance decision.
    // count > features.size() is an out-of-bounds read in the reference (:17-23 index features[0 .. count)); here it
    // means "all of them"
    if (count == 0 || count > features.size())
        count = features.size();
    if (count == 0)
        return {};
    // :17-23: indices 0 .. count-1 sorted by strength, descending, with std::sort. The sort runs on compact (strength,
    // index) records instead of indices compared through the 96-byte feature structs: std::sort's sequence of
    // comparisons and moves depends only on the comparator's answers, which are the same, so the permutation is the
    // reference's (ties included) -- and the records stay in cache.
    struct Ranked
    {
        float strength;
        uint32_t idx;
    };
    const bool compact = count <= 0xFFFFFFFFull;
    std::vector<size_t> by_strength(count);
    if (compact)
    {
        std::vector<Ranked> ranked(count);
        for (size_t i = 0; i < count; i++)
            ranked[i] = Ranked{features[i].strength, (uint32_t)i};
        std::sort(ranked.begin(), ranked.end(), [](const Ranked &a, const Ranked &b) { return a.strength > b.strength; });
        for (size_t i = 0; i < count; i++)
            by_strength[i] = ranked[i].idx;
    }
    else
    {
        for (size_t i = 0; i < count; i++)
            by_strength[i] = i;
        std::sort(by_strength.begin(), by_strength.end(),
                  [&features](size_t a, size_t b) { return features[a].strength > features[b].strength; });
    }

    std::vector<size_t> kept;
    kept.reserve(count);
    const double limit = spacing_pixels * spacing_pixels;
    const bool gridded = spacing_pixels > 0 && std::isfinite(spacing_pixels);
    // kept points bucketed by grid cell in a flat chained hash table. The cell is a little more than TWICE the
    // spacing (spacing = 0.4975 cells): the interval [x - spacing, x + spacing] then touches the point's own cell and
    // at most one neighbour per axis -- the lower one when the point lies in the lower half of its cell, else the upper
    // one -- with 0.0025 cells to spare on either side, far more than the rounding of the quotient x / cell
    // (computed as x * (1 / cell); |x / cell| < 1e9: error below 3e-7 cells). Four cells hold every kept point that can reject the candidate.
    const double cell = spacing_pixels * 2.01, inv_cell = 1.0 / cell;
    size_t table_size = 64;
    while (table_size < 2 * count)
        table_size <<= 1;
    std::vector<int32_t> head(table_size, -1);
    struct Entry
    {
        int64_t cx, cy;
        int32_t next;
        uint32_t pad;
        double x, y;
    };
    std::vector<Entry> entries;
    entries.reserve(count);
    auto slot_of = [table_size](int64_t cx, int64_t cy) {
        return static_cast<size_t>((static_cast<uint64_t>(cx) * 0x9E3779B97F4A7C15ull) ^
                                   (static_cast<uint64_t>(cy) * 0xC2B2AE3D27D4EB4Full)) >> 20 & (table_size - 1);
    };
    auto too_close = [&](size_t idx, size_t other) {
        const double dx = features[idx].location.x() - features[other].location.x();
        const double dy = features[idx].location.y() - features[other].location.y();
        return !(dx * dx + dy * dy > limit);
    };
    std::vector<size_t> unbucketed; // kept points with non-finite coordinates (compared against everything)
    for (size_t idx : by_strength)
    {
        const double x = features[idx].location.x(), y = features[idx].location.y();
        const double qx = x * inv_cell, qy = y * inv_cell; // within 3e-7 cells of x / cell for |x / cell| < 1e9
        const bool finite = gridded && std::isfinite(x) && std::isfinite(y) && std::abs(qx) < 1e9 && std::abs(qy) < 1e9;
        bool keep = true;
        int64_t cx = 0, cy = 0;
        if (finite)
        {
            const double fx = std::floor(qx), fy = std::floor(qy);
            cx = (int64_t)fx, cy = (int64_t)fy;
            if (!kept.empty())
            {
                const int64_t nx = qx - fx < 0.5 ? cx - 1 : cx + 1, ny = qy - fy < 0.5 ? cy - 1 : cy + 1;
                const int64_t gxs[2] = {cx, nx}, gys[2] = {cy, ny};
                for (int a = 0; a < 2 && keep; a++)
                    for (int b = 0; b < 2 && keep; b++)
                        for (int32_t e = head[slot_of(gxs[a], gys[b])]; e >= 0; e = entries[e].next)
                        {
                            const Entry &en = entries[e];
                            if (en.cx != gxs[a] || en.cy != gys[b])
                                continue;
                            // the same subtraction, squares and sum as too_close(): (x - x') and (y - y') on the
                            // coordinates kept next to the key (no trip to the feature structs)
                            const double dx = x - en.x, dy = y - en.y;
                            if (!(dx * dx + dy * dy > limit))
                            {
                                keep = false;
                                break;
                            }
                        }
                for (size_t other : unbucketed)
                    if (keep && too_close(idx, other))
                        keep = false;
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
        if (!keep)
            continue;
        if (finite)
        {
            const size_t sl = slot_of(cx, cy);
            entries.push_back(Entry{cx, cy, head[sl], 0u, x, y});
            head[sl] = (int32_t)(entries.size() - 1);
        }
        else
            unbucketed.push_back(idx);
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
