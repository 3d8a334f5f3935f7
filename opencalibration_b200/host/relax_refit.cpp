// Batched form of the "maximum likelihood" loop of RelaxGroup::finalize (reference src/relax/relax_group.cpp:156-165),
// the second caller of homography_model::fitInliers / evaluate: after the camera models changed, every edge of the
// measurement graph recomputes its homography from its previous inliers with
//     for (int i = 0; i < 3; i++) { h.fitInliers(correspondences, inliers); h.evaluate(correspondences, inliers); }
// Here all edges of a batch advance together: the correspondences are bound once (ocb_corr_bind_batch), and every round
// is ONE request table of OCB_REQ_REFIT_EVALUATE entries (all-inlier DLT refit on the device, K3, followed by
// Model::evaluate, K2): one launch and one copy each way per round instead of two GPU round trips per edge and round.
// Each job ends exactly where the reference's loop would: same matrices, same inlier vectors, same score.
#include "models_detail.hpp"

#include <cstring>
#include <stdexcept>

namespace ocb_host
{
void refit_evaluate_batch(std::vector<RefitJob> &jobs, int rounds)
{
    using namespace opencalibration;
    const size_t n_jobs = jobs.size();
    if (n_jobs == 0 || rounds <= 0)
        return;
    struct State
    {
        std::vector<uint32_t> in_bits, out_bits;
        double m18[18];
        double score = 0;
        uint32_t count = 0;
        bool live = false;
    };
    std::vector<State> st(n_jobs);
    std::vector<ocb_corr_set> sets(n_jobs);
    for (size_t j = 0; j < n_jobs; j++)
    {
        RefitJob &job = jobs[j];
        if (!job.matches || !job.model || !job.inliers)
            throw std::invalid_argument("refit_evaluate_batch: null job member");
        const size_t n = job.matches->size();
        if (job.inliers->size() != n)
            throw std::invalid_argument("refit_evaluate_batch: inliers and correspondences differ in length");
        // an edge without correspondences: fitInliers on nothing leaves NaNs in the reference too (the LU of a 1 x 9
        // system); evaluate of nothing scores 0. Nothing to send to the device.
        st[j].live = n > 0;
        sets[j] = ocb_corr_set{st[j].live ? detail::corr_data(*job.matches) : nullptr, n, nullptr};
        if (!st[j].live)
        {
            job.model->fitInliers(*job.matches, *job.inliers);
            job.score = 0;
            continue;
        }
        st[j].in_bits.assign((n + 31) / 32, 0u);
        st[j].out_bits.assign((n + 31) / 32, 0u);
        for (size_t k = 0; k < n; k++)
            if ((*job.inliers)[k])
                st[j].in_bits[k >> 5] |= 1u << (k & 31);
    }
    detail::gpu_check(ocb_corr_bind_batch(sets.data(), sets.size()), "ocb_corr_bind_batch");
    std::vector<ocb_score_request> requests;
    for (int r = 0; r < rounds; r++)
    {
        requests.clear();
        for (size_t j = 0; j < n_jobs; j++)
        {
            if (!st[j].live)
                continue;
            ocb_score_request q;
            std::memset(&q, 0, sizeof q);
            q.set = (uint32_t)j;
            q.kind = OCB_MODEL_HOMOGRAPHY;
            q.mode = OCB_REQ_REFIT_EVALUATE;
            q.h = 1;
            q.thr = jobs[j].model->inlier_threshold;
            q.refit_bits = st[j].in_bits.data();
            q.models_out = st[j].m18;
            q.score = &st[j].score;
            q.count = &st[j].count;
            q.inlier_bits = st[j].out_bits.data();
            requests.push_back(q);
        }
        if (requests.empty())
            break;
        detail::gpu_check(ocb_score_requests(requests.data(), requests.size()), "ocb_score_requests");
        for (size_t j = 0; j < n_jobs; j++)
            if (st[j].live)
                st[j].in_bits.swap(st[j].out_bits); // evaluate's inliers are the next round's fitInliers input
    }
    ocb_corr_unbind();
    for (size_t j = 0; j < n_jobs; j++)
    {
        if (!st[j].live)
            continue;
        RefitJob &job = jobs[j];
        detail::unpack_model(st[j].m18, *job.model);
        const size_t n = job.matches->size();
        for (size_t k = 0; k < n; k++)
            (*job.inliers)[k] = (st[j].in_bits[k >> 5] >> (k & 31)) & 1u;
        job.score = st[j].score;
    }
}
} // namespace ocb_host
