// Guided matcher of the dense stage: the Hamming site of densifyMesh (reference src/dense/dense_stereo.cpp:244-281).
//
// In the reference this is not a function of its own: for every dense source feature and every candidate image the
// KD-tree radius search `nearby` (:244-246) is followed inline by a top-2 scan over descriptor distances (:251-273)
// and the acceptance rule (:275-276). The scan is the data-parallel part; it runs here as ONE submission of all the
// (source feature, candidate image) visits against one candidate image -- K4 through ocb_match_lists. The KD-tree
// search stays the caller's (it is the reference's own jk-tree code); the acceptance rule is evaluated here in
// double, on distance = hamming * (1.0 / 486) exactly as descriptor_distance (:56-59) produces it.
#include "guided_match.hpp"

#include "models_detail.hpp"

#include <algorithm>
#include <cstring>
#include <limits>
#include <stdexcept>

#include <ocb.h>

namespace ocb_host
{
namespace
{
constexpr double RATIO_THRESHOLD = 0.85;                  // src/dense/dense_stereo.cpp:51
constexpr double MAX_ABSOLUTE_DESCRIPTOR_DISTANCE = 0.35; // :53
} // namespace

std::vector<GuidedMatch> match_features_guided(const std::vector<opencalibration::feature_2d> &source,
                                               const std::vector<opencalibration::feature_2d> &candidates,
                                               const GuidedLists &lists)
{
    using opencalibration::feature_2d;
    const size_t n_lists = lists.query_feature.size();
    if (lists.begin.size() != n_lists + 1)
        throw std::invalid_argument("GuidedLists::begin must have one entry per list plus one");
    std::vector<GuidedMatch> result;
    if (n_lists == 0)
        return result;
    if (lists.begin[0] != 0 || lists.begin[n_lists] != lists.nearby.size())
        throw std::invalid_argument("GuidedLists::begin must start at 0 and end at nearby.size()");

    // query rows: one packed row per list; candidate rows: the contiguous feature range the lists touch (the dense
    // features of an image are a contiguous tail of its feature vector, dense_stereo.cpp:127-131)
    std::vector<uint64_t> q_rows(n_lists * OCB_ROW_WORDS);
    std::vector<uint32_t> list_query(n_lists);
    for (size_t l = 0; l < n_lists; l++)
    {
        std::memcpy(&q_rows[l * OCB_ROW_WORDS], static_cast<const void *>(&source.at(lists.query_feature[l]).descriptor),
                    OCB_ROW_BYTES);
        list_query[l] = (uint32_t)l;
    }
    size_t lo = std::numeric_limits<size_t>::max(), hi = 0;
    for (size_t f : lists.nearby)
        lo = std::min(lo, f), hi = std::max(hi, f);
    std::vector<uint64_t> c_rows;
    std::vector<uint32_t> list_candidates(lists.nearby.size());
    if (!lists.nearby.empty())
    {
        if (hi >= candidates.size())
            throw std::out_of_range("GuidedLists::nearby names a feature past the end of the candidate image");
        c_rows.resize((hi - lo + 1) * OCB_ROW_WORDS);
        for (size_t f = lo; f <= hi; f++)
            std::memcpy(&c_rows[(f - lo) * OCB_ROW_WORDS], static_cast<const void *>(&candidates[f].descriptor),
                        OCB_ROW_BYTES);
        for (size_t k = 0; k < lists.nearby.size(); k++)
            list_candidates[k] = (uint32_t)(lists.nearby[k] - lo);
    }
    std::vector<uint64_t> begin(lists.begin.begin(), lists.begin.end());
    std::vector<ocb_top2> top(n_lists);
    detail::gpu_check(ocb_match_lists(q_rows.data(), n_lists, c_rows.data(), c_rows.size() / OCB_ROW_WORDS,
                                      list_query.data(), begin.data(), list_candidates.data(), n_lists, top.data()),
                      "ocb_match_lists");

    const double inf = std::numeric_limits<double>::infinity();
    for (size_t l = 0; l < n_lists; l++)
    {
        const size_t len = lists.begin[l + 1] - lists.begin[l];
        if (len == 0) // :248-249
            continue;
        const double best = top[l].best_d == OCB_DIST_INF ? inf : top[l].best_d * (1.0 / OCB_DESCRIPTOR_BITS);
        const double second = top[l].second_d == OCB_DIST_INF ? inf : top[l].second_d * (1.0 / OCB_DESCRIPTOR_BITS);
        const bool good = len >= 2 ? best < RATIO_THRESHOLD * second : best < MAX_ABSOLUTE_DESCRIPTOR_DISTANCE; // :275-276
        if (good)
            result.push_back(GuidedMatch{l, lists.query_feature[l], lists.nearby[lists.begin[l] + top[l].best_k], best, second});
    }
    return result;
}

} // namespace ocb_host
