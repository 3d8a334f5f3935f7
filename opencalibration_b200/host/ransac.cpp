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
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <omp.h>
#include <numeric>
#include <random>
#include <stdexcept>
#include <string>

namespace ocb_host
{
namespace
{
thread_local RansacStats t_stats;
// default on; OCB_RANSAC_DEVICE_FIT=0 in the environment selects host fits (A/B timing without rebuilding)
std::atomic<int> g_device_fit{[] {
    const char *e = std::getenv("OCB_RANSAC_DEVICE_FIT");
    return (e && e[0] == '0') ? 0 : 1;
}()};

template <size_t K> struct HypothesisStream
{
    // ransac.cpp:72-158: PROSAC ordering, seeded engine, shuffled evaluation order, sampling lambdas
    const std::vector<opencalibration::correspondence> &matches;
    bool has_quality = false;
    std::vector<size_t> sorted_idx;
    std::vector<size_t> eval_order;
    std::default_random_engine generator{42};
    size_t prosac_n;

    explicit HypothesisStream(const std::vector<opencalibration::correspondence> &m, const uint32_t *quality_order = nullptr)
        : matches(m)
    {
        for (const auto &c : matches)
            if (c.quality != 0)
            {
                has_quality = true;
                break;
            }
        if (has_quality && quality_order)
            sorted_idx.assign(quality_order, quality_order + matches.size()); // the same permutation, computed on the device
        else if (has_quality)
        {
            // ransac.cpp:83-90: indices sorted by quality, ascending, with std::sort. Sorted as compact (quality, index)
            // records: the comparator's answers, hence the permutation (ties included), are the reference's, without a
            // trip through the 56-byte correspondences for every comparison.
            struct Ranked
            {
                double quality;
                size_t idx;
            };
            std::vector<Ranked> ranked(matches.size());
            for (size_t i = 0; i < ranked.size(); i++)
                ranked[i] = Ranked{matches[i].quality, i};
            std::sort(ranked.begin(), ranked.end(),
                      [](const Ranked &a, const Ranked &b) { return a.quality < b.quality; });
            sorted_idx.resize(matches.size());
            for (size_t i = 0; i < ranked.size(); i++)
                sorted_idx[i] = ranked[i].idx;
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
void set_ransac_device_fit(bool on)
{
    g_device_fit.store(on ? 1 : 0);
}
bool ransac_device_fit()
{
    return g_device_fit.load() != 0;
}
} // namespace ocb_host

namespace opencalibration
{
using namespace ocb_host;
using namespace ocb_host::detail;

// One RANSAC run as a resumable state machine: advance() executes the reference's control flow
// (src/model_inliers/ransac.cpp:162-256) until it needs a GPU result -- the scores of a batch of hypotheses, the
// residuals of a would-be improver, or an evaluate() of the model in hand -- and returns what it needs; the driver
// (one run: ransac() below; many runs in lock step: ransac_batch()) obtains it and calls advance() again. Both
// drivers therefore execute the same host logic; only how requests reach the GPU differs.
template <typename Model> class RansacRun
{
  public:
    static constexpr size_t K = Model::MINIMUM_POINTS;
    enum class Need
    {
        DONE,
        SCORE_BATCH, // request_models(): want x 18, evaluation order -> batch_score, batch_count
        FIT_SCORE_BATCH, // request_samples(): want x 4 indices -> fitted on the device into batch_m18 / batch_skip,
                         // then scored like SCORE_BATCH (homography only)
        RESIDUALS,   // request_models(): 1 x 18                      -> residual[N]
        EVALUATE,    // request_models(): 1 x 18, index order         -> eval_score, eval_bits
        REFIT_EVALUATE // refit_bits (the current inliers): fitInliers on the device -> refit_m18, then like EVALUATE
                       // (homography only; drivers that cannot serve it leave device_refit off)
    };

    RansacRun(const std::vector<correspondence> &matches_, Model &model_, std::vector<bool> &inliers_,
              const uint32_t *quality_order = nullptr)
        : matches(matches_), model(model_), inliers(inliers_), N(matches_.size())
    {
        inliers.resize(N);
        std::fill(inliers.begin(), inliers.end(), false);
        if (N < K) // ransac.cpp:64-70
        {
            state = State::FINISHED;
            return;
        }
        stream.reset(new HypothesisStream<K>(matches, quality_order));
        order32.resize(N);
        for (size_t p = 0; p < N; p++)
            order32[p] = static_cast<uint32_t>(stream->eval_order[p]);
        kind = model_kind(model);
        thr = model.inlier_threshold;
        residual.resize(N);
        eval_bits.resize((N + 31) / 32);
        candidate_inliers.assign(N, false);
    }

    bool trivial() const { return N < K; }
    const uint32_t *order() const { return order32.data(); }
    const double *request_models() const { return request_m18; }
    size_t request_count() const { return request_h; }
    const uint32_t *request_samples() const { return batch_samples.data(); }
    double *fitted_models() { return batch_m18.data(); }
    uint8_t *fitted_degenerate() { return batch_skip.data(); }

    Need advance()
    {
        for (;;)
        {
            switch (state)
            {
            case State::FINISHED:
                return Need::DONE;

            case State::START_BATCH: {
                if (!(i < probability_iterations))
                {
                    // ransac.cpp:255-256: model = best_model; return model.evaluate(matches, inliers) / N
                    stats.iterations = i;
                    model = best_model;
                    request_evaluate();
                    state = State::AFTER_FINAL_EVAL;
                    return Need::EVALUATE;
                }
                const size_t want = std::min(batch, MAX_ITERATIONS - i);
                batch_skip.assign(want, 0);
                batch_m18.assign(want * 18, 0.0);
                batch_score.assign(want, 0.0);
                batch_count.assign(want, 0);
                batch_size = want;
                if constexpr (is_homography<Model>)
                {
                    if (device_fit)
                    {
                        // only the sample stream stays on the host (it is the reference's std:: RNG calls); the
                        // degeneracy test, the fit and the scoring of the whole batch run on the device
                        batch_samples.resize(want * K);
                        for (size_t k = 0; k < want; k++)
                        {
                            const std::array<size_t, K> sample = stream->next(i + k);
                            for (size_t j = 0; j < K; j++)
                                batch_samples[k * K + j] = static_cast<uint32_t>(sample[j]);
                        }
                        stats.gpu_calls++;
                        stats.scored += want;
                        b = 0;
                        request_m18 = nullptr;
                        request_h = want;
                        state = State::SCAN;
                        return Need::FIT_SCORE_BATCH;
                    }
                }
                batch_models.assign(want, model);
                for (size_t k = 0; k < want; k++)
                {
                    const std::array<size_t, K> sample = stream->next(i + k);
                    if constexpr (is_homography<Model>)
                    {
                        if (Model::checkSampleDegeneracy(matches, sample)) // ransac.cpp:173-177
                        {
                            batch_skip[k] = 1;
                            continue;
                        }
                    }
                    batch_models[k].fit(matches, sample); // ransac.cpp:179
                    pack_model(batch_models[k], &batch_m18[k * 18]);
                }
                stats.gpu_calls++;
                stats.scored += want;
                b = 0;
                request_m18 = batch_m18.data();
                request_h = want;
                state = State::SCAN;
                return Need::SCORE_BATCH;
            }

            case State::SCAN: {
                bool requested = false;
                for (; b < batch_size && i < probability_iterations; b++, i++)
                {
                    if (batch_skip[b])
                    {
                        stats.degenerate++;
                        continue;
                    }
                    if (!(batch_score[b] > best_score))
                        continue; // rejected early or not, nothing changes (ransac.cpp:204-207)
                    // would-be improver: replay ransac.cpp:183-203 on its residuals
                    if (batch_models.empty())
                        unpack_model(&batch_m18[b * 18], model); // fitted on the device
                    else
                        model = batch_models[b];
                    stats.gpu_calls++;
                    request_m18 = &batch_m18[b * 18];
                    request_h = 1;
                    requested = true;
                    break;
                }
                if (requested)
                {
                    state = State::AFTER_RESIDUALS;
                    return Need::RESIDUALS;
                }
                if (b >= batch_size)
                    batch = std::min<size_t>(batch * 2, 2048);
                state = State::START_BATCH;
                break;
            }

            case State::AFTER_RESIDUALS: {
                double score = 0;
                size_t checked = 0;
                bool rejected = false;
                std::fill(candidate_inliers.begin(), candidate_inliers.end(), false);
                for (size_t idx : stream->eval_order)
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
                    stats.rejected++;
                if (rejected || !(score > best_score)) // ransac.cpp:207
                {
                    next_hypothesis();
                    break;
                }
                stats.improvements++;
                best_model = model;
                best_score = score;
                inliers = candidate_inliers;
                if constexpr (is_fundamental<Model>) // ransac.cpp:213-222
                {
                    model.checkDegeneracy(matches, inliers);
                    request_evaluate();
                    state = State::AFTER_DEGEN_EVAL;
                    return Need::EVALUATE;
                }
                lo_round = 0;
                state = State::AFTER_LO_EVAL;
                return refit_and_evaluate(); // ransac.cpp:224
            }

            case State::AFTER_DEGEN_EVAL: {
                const double degen_score = take_evaluate();
                if (degen_score > best_score)
                {
                    best_model = model;
                    best_score = degen_score;
                }
                lo_round = 0;
                state = State::AFTER_LO_EVAL;
                return refit_and_evaluate();
            }

            case State::AFTER_LO_EVAL: {
                // ransac.cpp:225-245: keep refitting on the inliers of the last evaluate while the score improves
                if (pending_refit)
                {
                    unpack_model(refit_m18, model); // fitted on the device
                    pending_refit = false;
                }
                const double inlier_score = take_evaluate();
                const bool better = inlier_score > best_score;
                if (better)
                {
                    best_model = model;
                    best_score = inlier_score;
                }
                if (better && lo_round + 1 < MAX_INNER_ITERATIONS)
                {
                    lo_round++;
                    return refit_and_evaluate();
                }
                const double omega = best_score / N; // ransac.cpp:247-251
                const double omega_n = small_pow<(int)K>(omega);
                const double log_1m_omega_n = std::log(1 - omega_n);
                probability_iterations = std::max(
                    MIN_ITERATIONS, std::min(MAX_ITERATIONS, static_cast<size_t>(log_1m_p / log_1m_omega_n)));
                next_hypothesis();
                break;
            }

            case State::AFTER_FINAL_EVAL:
                result = take_evaluate() / N;
                state = State::FINISHED;
                return Need::DONE;
            }
        }
    }

    // ---- inputs / outputs of the run
    const std::vector<correspondence> &matches;
    Model &model;
    std::vector<bool> &inliers;
    const size_t N;
    int kind = 0;
    double thr = 0;
    double result = 0;
    RansacStats stats;
    // ---- result sinks the driver fills
    std::vector<double> batch_score, residual;
    std::vector<uint32_t> batch_count, eval_bits;
    double eval_score = 0;
    uint32_t eval_count = 0;
    // ---- all-inlier refits on the device (set by drivers that serve Need::REFIT_EVALUATE)
    bool device_refit = false;
    std::vector<uint32_t> refit_bits; // the inliers to refit to, index order
    double refit_m18[18];             // the refitted model

  private:
    enum class State
    {
        START_BATCH,
        SCAN,
        AFTER_RESIDUALS,
        AFTER_DEGEN_EVAL,
        AFTER_LO_EVAL,
        AFTER_FINAL_EVAL,
        FINISHED
    };
    static constexpr size_t MIN_ITERATIONS = 20, MAX_ITERATIONS = 10000, MAX_INNER_ITERATIONS = 5;
    const double log_1m_p = std::log(1 - 0.999); // PROBABILITY, ransac.cpp:60

    void next_hypothesis()
    {
        b++, i++;
        state = State::SCAN;
    }
    // model.fitInliers(matches, inliers) followed by the evaluate request (ransac.cpp:224-226, :233-235)
    Need refit_and_evaluate()
    {
        if constexpr (is_homography<Model>)
        {
            if (device_refit)
            {
                refit_bits.assign((N + 31) / 32, 0u);
                for (size_t k = 0; k < N; k++)
                    if (inliers[k])
                        refit_bits[k >> 5] |= 1u << (k & 31);
                pending_refit = true;
                request_m18 = nullptr;
                request_h = 1;
                stats.gpu_calls++;
                return Need::REFIT_EVALUATE;
            }
        }
        model.fitInliers(matches, inliers);
        request_evaluate();
        return Need::EVALUATE;
    }
    void request_evaluate()
    {
        pack_model(model, eval_m18);
        request_m18 = eval_m18;
        request_h = 1;
        stats.gpu_calls++;
    }
    // Model::evaluate's effect on `inliers` (homography_model.cpp:101-116) from the returned bit mask
    double take_evaluate()
    {
        inliers.resize(N);
        for (size_t k = 0; k < N; k++)
            inliers[k] = (eval_bits[k >> 5] >> (k & 31)) & 1u;
        return eval_score;
    }

    State state = State::START_BATCH;
    std::unique_ptr<HypothesisStream<K>> stream;
    std::vector<uint32_t> order32;
    Model best_model{};
    double best_score = 0;
    size_t probability_iterations = MAX_ITERATIONS;
    size_t i = 0;      // the reference's loop counter
    size_t batch = 32; // grows geometrically: the adaptive stop usually fires within the first batches
    size_t b = 0;      // cursor in the current batch
    size_t lo_round = 0;
    bool pending_refit = false;
    const bool device_fit = is_homography<Model> && ransac_device_fit();
    size_t batch_size = 0;
    std::vector<Model> batch_models; // host fits only; empty when the batch was fitted on the device
    std::vector<uint8_t> batch_skip;
    std::vector<uint32_t> batch_samples;
    std::vector<double> batch_m18;
    std::vector<bool> candidate_inliers;
    double eval_m18[18];
    const double *request_m18 = nullptr;
    size_t request_h = 0;
};

template <typename Model>
double ransac(const std::vector<correspondence> &matches, Model &model, std::vector<bool> &inliers)
{
    RansacRun<Model> run(matches, model, inliers);
    if (run.trivial())
    {
        t_stats = run.stats;
        return 0;
    }
    run.device_refit = ransac_device_fit(); // only the homography run ever asks for it
    // the correspondences stay on the GPU for the whole run: every request below sends only its models
    const BoundCorrespondences resident(matches, run.order());
    using Need = typename RansacRun<Model>::Need;
    for (Need need = run.advance(); need != Need::DONE; need = run.advance())
    {
        switch (need)
        {
        case Need::SCORE_BATCH:
            gpu_score_in_order(run.kind, run.request_models(), run.request_count(), matches, run.thr, run.order(),
                               run.batch_score.data(), run.batch_count.data());
            break;
        case Need::FIT_SCORE_BATCH:
            gpu_check(ocb_fit_score_bound(run.request_samples(), run.request_count(), run.thr, run.fitted_models(),
                                          run.fitted_degenerate(), run.batch_score.data(), run.batch_count.data()),
                      "ocb_fit_score_bound");
            break;
        case Need::RESIDUALS:
            gpu_residuals(run.kind, run.request_models(), matches, run.residual.data());
            break;
        case Need::EVALUATE:
            gpu_evaluate_bits(run.kind, run.request_models(), run.thr, matches, &run.eval_score, &run.eval_count,
                              run.eval_bits.data());
            break;
        case Need::REFIT_EVALUATE:
            gpu_check(ocb_refit_evaluate_bound(run.refit_bits.data(), run.thr, run.refit_m18, &run.eval_score,
                                               &run.eval_count, run.eval_bits.data()),
                      "ocb_refit_evaluate_bound");
            break;
        case Need::DONE:
            break;
        }
    }
    t_stats = run.stats;
    return run.result;
}

} // namespace opencalibration

namespace ocb_host
{
// Many runs advanced in lock step: each round every unfinished run contributes one request, and ONE
// ocb_score_requests call (one kernel launch, one copy each way) serves them all. The host-side logic of the runs
// (sampling, fits, SPRT replay, local optimisation) executes on `threads` OpenMP workers between rounds.
template <typename Model> void ransac_batch(std::vector<RansacJob<Model>> &jobs, int threads, const CorrBinder *binder)
{
    using namespace opencalibration;
    using Run = opencalibration::RansacRun<Model>;
    using Need = typename Run::Need;
    const size_t n_jobs = jobs.size();
    if (threads <= 0)
        threads = omp_get_num_procs();
    std::vector<std::unique_ptr<Run>> runs(n_jobs);
    std::string error;
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
    for (size_t j = 0; j < n_jobs; j++)
    {
        runs[j].reset(new Run(*jobs[j].matches, *jobs[j].model, *jobs[j].inliers, jobs[j].quality_order));
        runs[j]->device_refit = opencalibration::ransac_device_fit(); // only the homography run ever asks for it
    }
    std::vector<ocb_corr_set> sets(n_jobs);
    std::vector<size_t> active;
    for (size_t j = 0; j < n_jobs; j++)
    {
        const bool live = !runs[j]->trivial();
        sets[j] = ocb_corr_set{live ? detail::corr_data(*jobs[j].matches) : nullptr, live ? jobs[j].matches->size() : 0,
                               live ? runs[j]->order() : nullptr};
        if (live)
            active.push_back(j);
        else
            jobs[j].result = 0;
    }
    if (binder)
        (*binder)(sets);
    else
        detail::gpu_check(ocb_corr_bind_batch(sets.data(), sets.size()), "ocb_corr_bind_batch");
    std::vector<Need> need(n_jobs, Need::DONE);
    std::vector<ocb_score_request> requests;
    static const bool profile = std::getenv("OCB_RANSAC_PROFILE") != nullptr;
    double t_host = 0, t_gpu = 0;
    size_t rounds = 0, n_req[6] = {0, 0, 0, 0, 0, 0};
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto secs = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
        return std::chrono::duration<double>(b - a).count();
    };
    while (!active.empty())
    {
        const auto t0 = now();
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
        for (size_t a = 0; a < active.size(); a++)
        {
            try
            {
                need[active[a]] = runs[active[a]]->advance();
            }
            catch (const std::exception &e)
            {
#pragma omp critical(ocb_ransac_batch_error)
                error = e.what();
            }
        }
        if (!error.empty())
            throw std::runtime_error(error);
        requests.clear();
        std::vector<size_t> still;
        for (size_t j : active)
        {
            Run &r = *runs[j];
            if (need[j] == Need::DONE)
            {
                jobs[j].result = r.result;
                jobs[j].stats = r.stats;
                continue;
            }
            still.push_back(j);
            ocb_score_request q;
            std::memset(&q, 0, sizeof q);
            q.set = (uint32_t)j;
            q.kind = r.kind;
            q.h = (uint32_t)r.request_count();
            q.models = r.request_models();
            q.thr = r.thr;
            if (need[j] == Need::SCORE_BATCH)
                q.mode = OCB_REQ_SCORE_ORDERED, q.score = r.batch_score.data(), q.count = r.batch_count.data();
            else if (need[j] == Need::FIT_SCORE_BATCH)
                q.mode = OCB_REQ_FIT_SCORE_ORDERED, q.samples = r.request_samples(), q.models_out = r.fitted_models(),
                q.degenerate = r.fitted_degenerate(), q.score = r.batch_score.data(), q.count = r.batch_count.data();
            else if (need[j] == Need::REFIT_EVALUATE)
                q.mode = OCB_REQ_REFIT_EVALUATE, q.refit_bits = r.refit_bits.data(), q.models_out = r.refit_m18,
                q.score = &r.eval_score, q.count = &r.eval_count, q.inlier_bits = r.eval_bits.data();
            else if (need[j] == Need::RESIDUALS)
                q.mode = OCB_REQ_RESIDUALS, q.residuals = r.residual.data();
            else
                q.mode = OCB_REQ_EVALUATE, q.score = &r.eval_score, q.count = &r.eval_count,
                q.inlier_bits = r.eval_bits.data();
            requests.push_back(q);
        }
        const auto t1 = now();
        if (!requests.empty())
            detail::gpu_check(ocb_score_requests(requests.data(), requests.size()), "ocb_score_requests");
        if (profile)
        {
            t_host += secs(t0, t1), t_gpu += secs(t1, now()), rounds++;
            for (const ocb_score_request &q : requests)
                n_req[q.mode]++;
        }
        active.swap(still);
    }
    if (profile)
        std::fprintf(stderr,
                     "[ocb ransac_batch] jobs %zu threads %d rounds %zu host %.3f ms gpu %.3f ms requests: score %zu "
                     "evaluate %zu residuals %zu fit+score %zu refit+evaluate %zu\n",
                     n_jobs, threads, rounds, t_host * 1e3, t_gpu * 1e3, n_req[0], n_req[1], n_req[2], n_req[3], n_req[4]);
}
template void ransac_batch(std::vector<RansacJob<opencalibration::homography_model>> &, int, const CorrBinder *);
template void ransac_batch(std::vector<RansacJob<opencalibration::fundamental_matrix_model>> &, int, const CorrBinder *);
template void ransac_batch(std::vector<RansacJob<opencalibration::essential_matrix_model>> &, int, const CorrBinder *);
} // namespace ocb_host

namespace opencalibration
{
using namespace ocb_host;
using namespace ocb_host::detail;

template double ransac(const std::vector<correspondence> &, homography_model &, std::vector<bool> &);
template double ransac(const std::vector<correspondence> &, fundamental_matrix_model &, std::vector<bool> &);
template double ransac(const std::vector<correspondence> &, essential_matrix_model &, std::vector<bool> &);

void assembleInliers(const std::vector<feature_match> &matches, const std::vector<bool> &inliers,
                     const std::vector<feature_2d> &source_features, const std::vector<feature_2d> &dest_features,
                     std::vector<feature_match_denormalized> &inlier_list)
{
    // ransac.cpp:263-282
    inlier_list.reserve(std::count(inliers.begin(), inliers.end(), true));
    constexpr size_t AHEAD = 12; // the two feature records of a match are random 96-byte reads: ask for them early
    for (size_t i = 0; i < matches.size(); i++)
    {
        if (i + AHEAD < matches.size())
        {
            __builtin_prefetch(&source_features[matches[i + AHEAD].feature_index_1]);
            __builtin_prefetch(&dest_features[matches[i + AHEAD].feature_index_2]);
        }
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
