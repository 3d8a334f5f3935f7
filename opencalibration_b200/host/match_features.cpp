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
    // query is answered from a uniform grid of kept points (cell = spacing) instead of the reference's KD-tree,
    // which yields the same minimum-distance decision.
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
    // kept points bucketed by grid cell in a flat chained hash table. The cell is a little LARGER than the spacing: two
    // points no farther apart than `spacing` then differ by less than 0.999 cells per axis, so even with the quotient
    // x / cell rounded (|x / cell| < 1e9: error below 2.3e-7 cells) their cell indices differ by at most one, and the
    // 3 x 3 neighbourhood holds every kept point that can reject the candidate
    const double cell = spacing_pixels * 1.001;
    size_t table_size = 64;
    while (table_size < 2 * count)
        table_size <<= 1;
    std::vector<int32_t> head(table_size, -1);
    struct Entry
    {
        uint64_t key;
        int32_t next;
        uint32_t pad;
        double x, y;
    };
    std::vector<Entry> entries;
    entries.reserve(count);
    auto cell_key = [](int64_t cx, int64_t cy) {
        return (static_cast<uint64_t>(static_cast<uint32_t>(cx)) << 32) | static_cast<uint32_t>(cy);
    };
    auto slot_of = [table_size](uint64_t key) {
        key ^= key >> 33;
        key *= 0xff51afd7ed558ccdull;
        key ^= key >> 33;
        return static_cast<size_t>(key) & (table_size - 1);
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
        const bool finite = gridded && std::isfinite(x) && std::isfinite(y) && std::abs(x / cell) < 1e9 &&
                            std::abs(y / cell) < 1e9;
        bool keep = true;
        int64_t cx = 0, cy = 0;
        if (finite)
            cx = (int64_t)std::floor(x / cell), cy = (int64_t)std::floor(y / cell);
        if (!kept.empty())
        {
            if (finite)
            {
                for (int64_t gx = cx - 1; gx <= cx + 1 && keep; gx++)
                    for (int64_t gy = cy - 1; gy <= cy + 1 && keep; gy++)
                    {
                        const uint64_t key = cell_key(gx, gy);
                        for (int32_t e = head[slot_of(key)]; e >= 0; e = entries[e].next)
                        {
                            if (entries[e].key != key)
                                continue;
                            // the same subtraction, squares and sum as too_close(): (x - x') and (y - y') on the
                            // coordinates kept next to the key (no trip to the feature structs)
                            const double dx = x - entries[e].x, dy = y - entries[e].y;
                            if (!(dx * dx + dy * dy > limit))
                            {
                                keep = false;
                                break;
                            }
                        }
                    }
                for (size_t other : unbucketed)
                    if (keep && too_close(idx, other))
                        keep = false;
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
        {
            const uint64_t key = cell_key(cx, cy);
            const size_t sl = slot_of(key);
            entries.push_back(Entry{key, head[sl], 0u, x, y});
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
