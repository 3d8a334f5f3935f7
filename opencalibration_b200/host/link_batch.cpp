// Batched LinkStage runner -- see link_batch.hpp. Reference: src/pipeline/link_stage.cpp:41-117.
#include "link_batch.hpp"
#include "models_detail.hpp"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstring>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <omp.h>
#include <stdexcept>

namespace
{
using namespace opencalibration;
using clock_type = std::chrono::steady_clock;

double since(clock_type::time_point t0)
{
    return std::chrono::duration<double>(clock_type::now() - t0).count();
}

std::atomic<uint64_t> g_next_set_id{0x4c4e4b0000000000ull}; // ids of descriptor sets owned by link_pairs calls

// the tail of one LinkStage closure after the match (link_stage.cpp:86-108)
void finish_pair(const ocb_host::LinkImage &img, const ocb_host::LinkImage &near_image, camera_relations &relations,
                 std::vector<feature_match> &&coarse_matches, size_t *n_inliers)
{
    std::vector<correspondence> coarse_correspondences =
        distort_keypoints(*img.features, *near_image.features, coarse_matches, img.model, near_image.model);
    homography_model h;
    std::vector<bool> coarse_inliers;
    ransac(coarse_correspondences, h, coarse_inliers);
    relations.ransac_relation = h.homography;
    relations.relationType = camera_relations::RelationType::HOMOGRAPHY;
    const bool can_decompose = h.decompose(coarse_correspondences, coarse_inliers, relations.relative_poses);
    const size_t num_coarse_inliers = std::count(coarse_inliers.begin(), coarse_inliers.end(), true);
    *n_inliers = num_coarse_inliers;
    if (can_decompose && num_coarse_inliers > h.MINIMUM_POINTS * 1.5)
    {
        relations.matches = std::move(coarse_matches);
        assembleInliers(relations.matches, coarse_inliers, *img.features, *near_image.features,
                        relations.inlier_matches);
    }
}
} // namespace

namespace ocb_host
{
std::vector<camera_relations> link_pairs(const std::vector<LinkImage> &images, const std::vector<LinkPair> &pairs,
                                         const LinkOptions &options, LinkStats *stats)
{
    const auto t_begin = clock_type::now();
    const int threads = options.threads > 0 ? options.threads : omp_get_num_procs();
    const size_t n_img = images.size(), n_pairs = pairs.size();
    for (const LinkPair &p : pairs)
        if (p.image_1 >= n_img || p.image_2 >= n_img)
            throw std::invalid_argument("link_pairs: pair references an unknown image");
    for (const LinkImage &im : images)
        if (!im.features)
            throw std::invalid_argument("link_pairs: image without features");

    // ---- per image, once: the subsample every closure of that image would compute (link_stage.cpp:63-65,80-81) and
    // the upload of those rows. Only images that occur in a pair are touched.
    std::vector<char> used(n_img, 0);
    for (const LinkPair &p : pairs)
        used[p.image_1] = used[p.image_2] = 1;
    std::vector<std::vector<size_t>> indices(n_img);
    const uint64_t id_base = g_next_set_id.fetch_add(n_img + 1);
    std::string error;
    // every worker (and the submission thread) runs on the process's default device: the one of its first ocb_init
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
    for (size_t i = 0; i < n_img; i++)
    {
        if (!used[i])
            continue;
        try
        {
            indices[i] = spatially_subsample_feature_indices(*images[i].features, options.coarse_spacing_pixels,
                                                             images[i].num_sparse_features);
        }
        catch (const std::exception &e)
        {
#pragma omp critical(ocb_link_error)
            error = e.what();
        }
    }
    // upload: one batched registration (one device allocation, pipelined gather + copy) per worker, rows taken
    // straight from the feature vectors through the subsample indices
    std::vector<size_t> used_ids;
    for (size_t i = 0; i < n_img; i++)
        if (used[i])
            used_ids.push_back(i);
    std::vector<char> registered(n_img, 0);
    const int groups = (int)std::min<size_t>((size_t)std::min(threads, 8), std::max<size_t>(used_ids.size(), 1));
    if (error.empty())
    {
#pragma omp parallel for schedule(static, 1) num_threads(groups)
        for (int g = 0; g < groups; g++)
        {
            std::vector<ocb_set_source> src;
            for (size_t u = (size_t)g; u < used_ids.size(); u += (size_t)groups)
            {
                const size_t i = used_ids[u];
                const std::vector<feature_2d> &f = *images[i].features;
                src.push_back(ocb_set_source{id_base + i, f.empty() ? nullptr : static_cast<const void *>(&f[0].descriptor),
                                             sizeof(feature_2d), indices[i].data(), indices[i].size()});
            }
            const int rc = ocb_register_descriptors_batch(src.data(), src.size());
#pragma omp critical(ocb_link_error)
            {
                if (rc)
                    error = std::string("ocb_register_descriptors_batch: ") + ocb_last_error();
                else
                    for (const ocb_set_source &sset : src)
                        registered[sset.set_id - id_base] = 1;
            }
        }
    }
    auto release_sets = [&]() {
        for (size_t i = 0; i < n_img; i++)
            if (registered[i])
                ocb_unregister_descriptors(id_base + i);
    };
    if (!error.empty())
    {
        release_sets();
        throw std::runtime_error(error);
    }
    LinkStats st;
    st.seconds_subsample_upload = since(t_begin);

    // ---- submissions: one long-lived thread keeps the GPU matching chunk k+1 (into the other of two page-locked
    // result buffers) while the OpenMP workers finish chunk k
    const size_t per = std::max<size_t>(1, options.pairs_per_submission);
    const size_t n_chunks = (n_pairs + per - 1) / per;
    size_t max_rows = 1;
    for (size_t c = 0; c < n_chunks; c++)
    {
        size_t rows = 0;
        for (size_t p = c * per; p < std::min(n_pairs, (c + 1) * per); p++)
            rows += indices[pairs[p].image_1].size();
        max_rows = std::max(max_rows, rows);
    }
    struct Slot
    {
        ocb_top2 *top = nullptr;
        std::vector<uint64_t> offsets;
        double gpu_seconds = 0;
        bool full = false;
    } slot[2];
    for (Slot &sl : slot)
    {
        sl.top = static_cast<ocb_top2 *>(ocb_host_alloc(max_rows * sizeof(ocb_top2)));
        if (!sl.top)
        {
            for (Slot &o : slot)
                ocb_host_free(o.top);
            release_sets();
            throw std::runtime_error(std::string("ocb_host_alloc: ") + ocb_last_error());
        }
    }
    std::mutex mu;
    std::condition_variable cv;
    std::string producer_error;
    bool stop = false;
    std::thread producer([&]() {
        for (size_t c = 0; c < n_chunks; c++)
        {
            Slot &sl = slot[c & 1];
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return !sl.full || stop; });
                if (stop)
                    return;
            }
            const size_t begin = c * per, end = std::min(n_pairs, begin + per);
            std::vector<ocb_pair> sub(end - begin);
            sl.offsets.resize(sub.size());
            uint64_t total = 0;
            for (size_t p = begin; p < end; p++)
            {
                sub[p - begin] = ocb_pair{id_base + pairs[p].image_1, id_base + pairs[p].image_2};
                sl.offsets[p - begin] = total;
                total += indices[pairs[p].image_1].size();
            }
            const auto t0 = clock_type::now();
            const int rc = ocb_match_pairs(sub.data(), sub.size(), sl.top, sl.offsets.data());
            sl.gpu_seconds = since(t0);
            std::lock_guard<std::mutex> lk(mu);
            if (rc)
            {
                producer_error = std::string("ocb_match_pairs: ") + ocb_last_error();
                stop = true;
            }
            sl.full = true;
            cv.notify_all();
            if (rc)
                return;
        }
    });

    std::vector<camera_relations> relations(n_pairs);
    size_t total_matches = 0, total_inliers = 0;
    for (size_t c = 0; c < n_chunks && error.empty(); c++)
    {
        Slot &sl = slot[c & 1];
        {
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&] { return sl.full; });
            if (!producer_error.empty())
            {
                error = producer_error;
                break;
            }
        }
        st.seconds_match_gpu += sl.gpu_seconds;
        const size_t begin = c * per, end = std::min(n_pairs, begin + per);
        const auto t0 = clock_type::now();
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads) reduction(+ : total_matches, total_inliers)
        for (size_t p = begin; p < end; p++)
        {
            try
            {
                const LinkImage &img = images[pairs[p].image_1], &near_image = images[pairs[p].image_2];
                std::vector<feature_match> coarse_matches =
                    detail::matches_from_top2(indices[pairs[p].image_1], indices[pairs[p].image_2],
                                              sl.top + sl.offsets[p - begin], nullptr, nullptr);
                total_matches += coarse_matches.size();
                if (options.run_ransac)
                {
                    size_t inl = 0;
                    finish_pair(img, near_image, relations[p], std::move(coarse_matches), &inl);
                    total_inliers += inl;
                }
                else
                    relations[p].matches = std::move(coarse_matches);
            }
            catch (const std::exception &e)
            {
#pragma omp critical(ocb_link_error)
                error = e.what();
            }
        }
        st.seconds_tail += since(t0);
        std::lock_guard<std::mutex> lk(mu);
        sl.full = false;
        cv.notify_all();
    }
    {
        std::lock_guard<std::mutex> lk(mu);
        stop = true;
        cv.notify_all();
    }
    producer.join();
    for (Slot &sl : slot)
        ocb_host_free(sl.top);
    release_sets();
    if (!error.empty())
        throw std::runtime_error(error);
    for (const LinkPair &p : pairs)
        st.comparisons += indices[p.image_1].size() * indices[p.image_2].size();
    st.matches = total_matches;
    st.ransac_inliers = total_inliers;
    st.seconds_total = since(t_begin);
    if (stats)
        *stats = st;
    return relations;
}
} // namespace ocb_host
