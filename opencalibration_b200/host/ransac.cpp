// ransac<Model>() and assembleInliers() with the reference's signatures
// (reference include/opencalibration/model_inliers/ransac.hpp:15-20, src/model_inliers/ransac.cpp:54-282),
// re-organised for a GPU: hypotheses are generated and fitted in batches on the host, scored as a
// hypotheses x correspondences grid on the device (K2/K3 via ocb_score_models), and the reference's sequential
// accept / SPRT-reject / local-optimisation / adaptive-termination logic is then replayed over the batch.
//
// Why the replay gives the reference's result:
//   * the hypothesis stream (shuffle of eval_order, PROSAC / uniform sampling) is drawn with the same std::
//     entities in the same order and never depends on scores -- drawing a few samples
//     more than the reference would have consumed is unobservable (the engine is local to the call);
//   * the device returns the full MSAC sum of every hypothesis, accumulated sequentially in eval_order with
//     individually rounded IEEE operations, i.e. the value `score` has at ransac.cpp:204 when the hypothesis
//     was not rejected early;
//   * a hypothesis whose full score does not beat best_score changes nothing whether or not the SPRT test
//     would have cut it short (ransac.cpp:204-207), so the prefix test (:197-202) only has to be replayed for
//     would-be improvers: their per-correspondence residuals are fetched (ocb_residuals) and the reference's
//     loop is run on them verbatim, which also yields candidate_inliers.
#include "models_detail.hpp"

#include <algorithm>
#include <cmath>
#include <numeric>
#include <random>

namespace ocb_host
{
namespace
{
thread_local RansacStats t_stats;

template <size_t K> struct HypothesisStream
{
    // ransac.cpp:72-158: PROSAC ordering, seeded engine, shuffled evaluation order, sampling lambdas
    const std::vector<opencalibration::correspondence> &matches;
    bool has_quality = false;
    std::vector<size_t> sorted_idx;
    std::vector<size_t> eval_order;
    std::default_random_engine generator{42};
    size_t prosac_n;

    explicit HypothesisStream(const std::vector<opencalibration::correspondence> &m) : matches(m)
    {
        for (const auto &c : matches)
            if (c.quality != 0)
            {
                has_quality = true;
                break;
            }
        if (has_quality)
        {
            sorted_idx.resize(matches.size());
            std::iota(sorted_idx.begin(), sorted_idx.end(), 0);
            std::sort(sorted_idx.begin(), sorted_idx.end(),
                      [this](size_t a, size_t b) { return matches[a].quality < matches[b].quality; });
        }
        eval_order.resize(matches.size());
        std::iota(eval_order.begin(), eval_order.end(), 0);
        prosac_n = has_quality ? K : matches.size();
        std::shuffle(eval_order.begin(), eval_order.end(), generator);
    }

    size_t mapped(size_t i) const { return has_quality ? sorted_idx[i] : i; }

    std::array<size_t, K> uniform_sample(size_t pool)
    {
        std::array<size_t, K> indices;
        std::uniform_int_distribution<size_t> dist(0, pool - 1);
        for (size_t j = 0; j < K; j++)
        {
            size_t candidate;
            bool unique;
            do
            {
                candidate = dist(generator);
                unique = true;
                for (size_t k = 0; k < j; k++)
                    if (indices[k] == mapped(candidate))
                    {
                        unique = false;
                        break;
                    }
            } while (!unique);
            indices[j] = mapped(candidate);
        }
        return indices;
    }

    std::array<size_t, K> growth_sample(size_t pool)
    {
        std::array<size_t, K> indices;
        indices[0] = sorted_idx[pool - 1];
        std::uniform_int_distribution<size_t> dist(0, pool - 2);
        for (size_t j = 1; j < K; j++)
        {
            size_t candidate;
            bool unique;
            do
            {
                candidate = dist(generator);
                unique = true;
                for (size_t k = 0; k < j; k++)
                    if (indices[k] == sorted_idx[candidate])
                    {
                        unique = false;
                        break;
                    }
            } while (!unique);
            indices[j] = sorted_idx[candidate];
        }
        return indices;
    }

    // the sample of loop iteration i (ransac.cpp:164-171); must be called for i = 0, 1, 2, ... in order
    std::array<size_t, K> next(size_t i)
    {
        if (has_quality && prosac_n < matches.size() && i > 0 && i % 10 == 0)
            prosac_n++;
        if (has_quality && prosac_n < matches.size() && prosac_n > K)
            return growth_sample(prosac_n);
        return uniform_sample(has_quality ? prosac_n : matches.size());
    }
};

template <int N> double small_pow(double d); // ransac.cpp:32-51
template <> inline double small_pow<4>(double d)
{
    const double t = d * d;
    return t * t;
}
template <> inline double small_pow<5>(double d)
{
    const double t = d * d;
    return t * t * d;
}
template <> inline double small_pow<8>(double d)
{
    double t = d * d;
    t = t * t;
    return t * t;
}

template <typename Model> constexpr bool is_homography = false;
template <> constexpr bool is_homography<opencalibration::homography_model> = true;
template <typename Model> constexpr bool is_fundamental = false;
template <> constexpr bool is_fundamental<opencalibration::fundamental_matrix_model> = true;

} // namespace

RansacStats last_ransac_stats()
{
    return t_stats;
}
} // namespace ocb_host

namespace opencalibration
{
using namespace ocb_host;
using namespace ocb_host::detail;

template <typename Model>
double ransac(const std::vector<correspondence> &matches, Model &model, std::vector<bool> &inliers)
{
    constexpr size_t K = Model::MINIMUM_POINTS;
    const size_t MIN_ITERATIONS = 20;
    const size_t MAX_ITERATIONS = 10000;
    const size_t MAX_INNER_ITERATIONS = 5;
    const double PROBABILITY = 0.999;
    const double log_1m_p = std::log(1 - PROBABILITY);
    const size_t N = matches.size();
    RansacStats stats;

    inliers.resize(N);
    std::fill(inliers.begin(), inliers.end(), false);
    if (N < K)
    {
        t_stats = stats;
        return 0;
    }

    HypothesisStream<K> stream(matches);
    std::vector<uint32_t> order32(N);
    for (size_t p = 0; p < N; p++)
        order32[p] = static_cast<uint32_t>(stream.eval_order[p]);

    // the correspondences stay on the GPU for the whole run: every batch, residual fetch and evaluate() below
    // sends only its models
    const BoundCorrespondences resident(matches, order32.data());

    Model best_model{};
    double best_score = 0;
    size_t probability_iterations = MAX_ITERATIONS;
    const int kind = model_kind(model);
    const double thr = model.inlier_threshold;

    std::vector<Model> batch_models;
    std::vector<char> batch_skip;
    std::vector<double> batch_m18, batch_score, residual(N);
    std::vector<uint32_t> batch_count;
    std::vector<bool> candidate_inliers(N, false);

    size_t i = 0;          // the reference's loop counter
    size_t batch = 32;     // grows geometrically: the adaptive stop usually fires within the first batches
    while (i < probability_iterations)
    {
        const size_t want = std::min(batch, MAX_ITERATIONS - i);
        batch_models.assign(want, model);
        batch_skip.assign(want, 0);
        batch_m18.assign(want * 18, 0.0);
        for (size_t b = 0; b < want; b++)
        {
            const std::array<size_t, K> sample = stream.next(i + b);
            if constexpr (is_homography<Model>)
            {
                if (Model::checkSampleDegeneracy(matches, sample)) // ransac.cpp:173-177
                {
                    batch_skip[b] = 1;
                    continue;
                }
            }
            batch_models[b].fit(matches, sample); // ransac.cpp:179
            pack_model(batch_models[b], &batch_m18[b * 18]);
        }
        batch_score.assign(want, 0.0);
        batch_count.assign(want, 0);
        gpu_score_in_order(kind, batch_m18.data(), want, matches, thr, order32.data(), batch_score.data(),
                           batch_count.data());
        stats.gpu_calls++;
        stats.scored += want;

        for (size_t b = 0; b < want && i < probability_iterations; b++, i++)
        {
            if (batch_skip[b])
            {
                stats.degenerate++;
                continue;
            }
            if (!(batch_score[b] > best_score))
                continue; // rejected early or not, nothing changes (ransac.cpp:204-207)

            // would-be improver: replay ransac.cpp:183-203 on its residuals
            model = batch_models[b];
            gpu_residuals(kind, &batch_m18[b * 18], matches, residual.data());
            stats.gpu_calls++;
            double score = 0;
            size_t checked = 0;
            bool rejected = false;
            std::fill(candidate_inliers.begin(), candidate_inliers.end(), false);
            for (size_t idx : stream.eval_order)
            {
                const double e = residual[idx];
                if (e < model.inlier_threshold)
                {
                    candidate_inliers[idx] = true;
                    const double ratio = e / model.inlier_threshold;
                    score += 1.0 - ratio * ratio;
                }
                checked++;
                if (checked > 20 && best_score > 0 && score < best_score * static_cast<double>(checked) / N * 0.6)
                {
                    rejected = true;
                    break;
                }
            }
            if (rejected)
            {
                stats.rejected++;
                continue;
            }
            if (score > best_score) // ransac.cpp:207
            {
                stats.improvements++;
                best_model = model;
                best_score = score;
                inliers = candidate_inliers;

                if constexpr (is_fundamental<Model>) // ransac.cpp:213-222
                {
                    model.checkDegeneracy(matches, inliers);
                    const double degen_score = model.evaluate(matches, inliers);
                    if (degen_score > best_score)
                    {
                        best_model = model;
                        best_score = degen_score;
                    }
                }

                model.fitInliers(matches, inliers); // ransac.cpp:224-245
                double inlier_score = model.evaluate(matches, inliers);
                if (inlier_score > best_score)
                {
                    best_model = model;
                    best_score = inlier_score;
                    for (size_t j = 1; j < MAX_INNER_ITERATIONS; j++)
                    {
                        model.fitInliers(matches, inliers);
                        inlier_score = model.evaluate(matches, inliers);
                        if (inlier_score > best_score)
                        {
                            best_model = model;
                            best_score = inlier_score;
                        }
                        else
                        {
                            break;
                        }
                    }
                }

                const double omega = best_score / N; // ransac.cpp:247-251
                const double omega_n = small_pow<(int)K>(omega);
                const double log_1m_omega_n = std::log(1 - omega_n);
                probability_iterations =
                    std::max(MIN_ITERATIONS, std::min(MAX_ITERATIONS, static_cast<size_t>(log_1m_p / log_1m_omega_n)));
            }
        }
        batch = std::min<size_t>(batch * 2, 2048);
    }
    stats.iterations = i;
    t_stats = stats;

    model = best_model;
    return model.evaluate(matches, inliers) / N; // ransac.cpp:255-256
}

template double ransac(const std::vector<correspondence> &, homography_model &, std::vector<bool> &);
template double ransac(const std::vector<correspondence> &, fundamental_matrix_model &, std::vector<bool> &);
template double ransac(const std::vector<correspondence> &, essential_matrix_model &, std::vector<bool> &);

void assembleInliers(const std::vector<feature_match> &matches, const std::vector<bool> &inliers,
                     const std::vector<feature_2d> &source_features, const std::vector<feature_2d> &dest_features,
                     std::vector<feature_match_denormalized> &inlier_list)
{
    // ransac.cpp:263-282
    inlier_list.reserve(std::count(inliers.begin(), inliers.end(), true));
    for (size_t i = 0; i < matches.size(); i++)
    {
        if (!inliers[i])
            continue;
        feature_match_denormalized d;
        d.pixel_1 = source_features[matches[i].feature_index_1].location;
        d.pixel_2 = dest_features[matches[i].feature_index_2].location;
        d.feature_index_1 = matches[i].feature_index_1;
        d.feature_index_2 = matches[i].feature_index_2;
        d.match_index = i;
        inlier_list.push_back(d);
    }
}

} // namespace opencalibration
