// Batched LinkStage runner -- see link_batch.hpp. Reference: src/pipeline/link_stage.cpp:41-117.
#include "link_batch.hpp"
#include "models_detail.hpp"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstring>
#include <future>
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
            const std::vector<feature_2d> &f = *images[i].features;
            indices[i] = spatially_subsample_feature_indices(f, options.coarse_spacing_pixels,
                                                             images[i].num_sparse_features);
            std::vector<uint64_t> rows(indices[i].size() * OCB_ROW_WORDS);
            for (size_t k = 0; k < indices[i].size(); k++)
                std::memcpy(&rows[k * OCB_ROW_WORDS], static_cast<const void *>(&f[indices[i][k]].descriptor),
                            OCB_ROW_BYTES);
            detail::gpu_check(ocb_register_descriptors(id_base + i, rows.data(), indices[i].size()),
                              "ocb_register_descriptors");
        }
        catch (const std::exception &e)
        {
#pragma omp critical(ocb_link_error)
            error = e.what();
        }
    }
    auto release_sets = [&]() {
        for (size_t i = 0; i < n_img; i++)
            if (used[i])
                ocb_unregister_descriptors(id_base + i);
    };
    if (!error.empty())
    {
        release_sets();
        throw std::runtime_error(error);
    }
    LinkStats st;
    st.seconds_subsample_upload = since(t_begin);

    // ---- submissions: GPU matches chunk k+1 while the workers finish chunk k
    struct Chunk
    {
        size_t begin = 0, end = 0;
        std::vector<ocb_top2> top;
        std::vector<uint64_t> offsets;
        double gpu_seconds = 0;
    };
    const size_t per = std::max<size_t>(1, options.pairs_per_submission);
    auto match_chunk = [&](size_t begin) {
        Chunk c;
        c.begin = begin, c.end = std::min(n_pairs, begin + per);
        std::vector<ocb_pair> sub(c.end - c.begin);
        c.offsets.resize(sub.size());
        uint64_t total = 0;
        for (size_t p = c.begin; p < c.end; p++)
        {
            sub[p - c.begin] = ocb_pair{id_base + pairs[p].image_1, id_base + pairs[p].image_2};
            c.offsets[p - c.begin] = total;
            total += indices[pairs[p].image_1].size();
        }
        c.top.resize(total);
        const auto t0 = clock_type::now();
        detail::gpu_check(ocb_match_pairs(sub.data(), sub.size(), c.top.data(), c.offsets.data()), "ocb_match_pairs");
        c.gpu_seconds = since(t0);
        return c;
    };

    std::vector<camera_relations> relations(n_pairs);
    size_t total_matches = 0, total_inliers = 0;
    try
    {
        std::future<Chunk> next;
        if (n_pairs)
            next = std::async(std::launch::async, match_chunk, (size_t)0);
        for (size_t begin = 0; begin < n_pairs; begin += per)
        {
            Chunk c = next.get();
            st.seconds_match_gpu += c.gpu_seconds;
            if (c.end < n_pairs)
                next = std::async(std::launch::async, match_chunk, c.end);
            const auto t0 = clock_type::now();
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads) reduction(+ : total_matches, total_inliers)
            for (size_t p = c.begin; p < c.end; p++)
            {
                try
                {
                    const LinkImage &img = images[pairs[p].image_1], &near_image = images[pairs[p].image_2];
                    std::vector<feature_match> coarse_matches =
                        detail::matches_from_top2(indices[pairs[p].image_1], indices[pairs[p].image_2],
                                                  c.top.data() + c.offsets[p - c.begin], nullptr, nullptr);
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
            if (!error.empty())
            {
                if (next.valid())
                    next.wait();
                throw std::runtime_error(error);
            }
        }
    }
    catch (...)
    {
        release_sets();
        throw;
    }
    release_sets();
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
