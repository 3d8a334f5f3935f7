// See partition.hpp. Reference: include/opencalibration/types/hilbert.hpp:8-27, src/pipeline/link_stage.cpp:75-131.
#include "partition.hpp"

#include <ocb.h>

#include <algorithm>
#include <chrono>
#include <exception>
#include <numeric>
#include <omp.h>
#include <stdexcept>
#include <string>
#include <thread>
#include <utility>

namespace ocb_host
{
uint32_t hilbert_index(int order, int x, int y)
{
    uint32_t d = 0;
    for (int s = order / 2; s > 0; s /= 2)
    {
        const int rx = (x & s) ? 1 : 0, ry = (y & s) ? 1 : 0;
        d += (uint32_t)s * (uint32_t)s * (uint32_t)((3 * rx) ^ ry);
        if (ry == 0)
        {
            if (rx == 1)
            {
                x = s - 1 - x;
                y = s - 1 - y;
            }
            std::swap(x, y);
        }
    }
    return d;
}

std::vector<size_t> hilbert_order(const double *xy, size_t n)
{
    std::vector<size_t> order(n);
    std::iota(order.begin(), order.end(), (size_t)0);
    if (n == 0 || !xy)
        return order;
    constexpr int GRID = 1024;
    double lo[2] = {xy[0], xy[1]}, hi[2] = {xy[0], xy[1]};
    for (size_t i = 0; i < n; i++)
        for (int a = 0; a < 2; a++)
        {
            lo[a] = std::min(lo[a], xy[2 * i + a]);
            hi[a] = std::max(hi[a], xy[2 * i + a]);
        }
    const double span[2] = {std::max(hi[0] - lo[0], 1e-12), std::max(hi[1] - lo[1], 1e-12)};
    std::vector<uint32_t> key(n);
    for (size_t i = 0; i < n; i++)
    {
        int cell[2];
        for (int a = 0; a < 2; a++)
        {
            const double scaled = (xy[2 * i + a] - lo[a]) / span[a] * GRID;
            cell[a] = (int)std::min<long long>((long long)scaled, GRID - 1);
        }
        key[i] = hilbert_index(GRID, cell[0], cell[1]);
    }
    std::stable_sort(order.begin(), order.end(), [&key](size_t a, size_t b) { return key[a] < key[b]; });
    return order;
}

std::vector<PairShard> partition_pairs(const double *xy, size_t n_images, const std::vector<LinkPair> &pairs,
                                       size_t world)
{
    if (world == 0)
        throw std::invalid_argument("partition_pairs: world must be positive");
    for (const LinkPair &p : pairs)
        if (p.image_1 >= n_images || p.image_2 >= n_images)
            throw std::invalid_argument("partition_pairs: pair references an unknown image");
    const std::vector<size_t> order = hilbert_order(xy, n_images);
    std::vector<size_t> load(n_images, 0);
    for (const LinkPair &p : pairs)
        load[p.image_1]++;
    const size_t total = pairs.size();
    // cut the curve where the running number of sourced pairs crosses a multiple of total / world
    std::vector<size_t> owner(n_images, 0);
    size_t before = 0;
    for (size_t img : order)
    {
        owner[img] = std::min(world - 1, before * world / std::max<size_t>(total, 1));
        before += load[img];
    }
    std::vector<PairShard> shards(world);
    for (size_t i = 0; i < n_images; i++)
        shards[owner[i]].owned_images.push_back(i);
    std::vector<std::vector<char>> is_halo(world);
    for (size_t k = 0; k < pairs.size(); k++)
    {
        const size_t r = owner[pairs[k].image_1];
        shards[r].pair_ids.push_back(k);
        if (owner[pairs[k].image_2] != r)
        {
            if (is_halo[r].empty())
                is_halo[r].assign(n_images, 0);
            is_halo[r][pairs[k].image_2] = 1;
        }
    }
    for (size_t r = 0; r < world; r++)
        if (!is_halo[r].empty())
            for (size_t i = 0; i < n_images; i++)
                if (is_halo[r][i])
                    shards[r].halo_images.push_back(i);
    return shards;
}

std::vector<opencalibration::camera_relations> link_pairs_multi(const std::vector<LinkImage> &images,
                                                                const std::vector<LinkPair> &pairs, const double *xy,
                                                                int n_devices, const LinkOptions &options,
                                                                LinkStats *stats)
{
    if (n_devices <= 0)
        throw std::invalid_argument("link_pairs_multi: n_devices must be positive");
    if (n_devices > ocb_device_count())
        throw std::runtime_error("link_pairs_multi: fewer CUDA devices than requested (there is no CPU fallback)");
    if (n_devices == 1)
        return link_pairs(images, pairs, options, stats);
    const auto t0 = std::chrono::steady_clock::now();
    const std::vector<PairShard> shards = partition_pairs(xy, images.size(), pairs, (size_t)n_devices);
    std::vector<opencalibration::camera_relations> relations(pairs.size());
    std::vector<LinkStats> part_stats((size_t)n_devices);
    std::vector<std::exception_ptr> errors((size_t)n_devices);
    const int all_threads = options.threads > 0 ? options.threads : omp_get_num_procs();
    const int caller_device = ocb_current_device();
    std::vector<std::thread> workers;
    for (int d = 0; d < n_devices; d++)
        workers.emplace_back([&, d]() {
            try
            {
                const PairShard &sh = shards[(size_t)d];
                if (sh.pair_ids.empty())
                    return;
                if (ocb_init(d))
                    throw std::runtime_error(std::string("ocb_init: ") + ocb_last_error());
                // the part's images, renumbered: own images first, then the halo
                std::vector<size_t> local(images.size(), ~(size_t)0);
                std::vector<LinkImage> part_images;
                for (const std::vector<size_t> *list : {&sh.owned_images, &sh.halo_images})
                    for (size_t i : *list)
                    {
                        local[i] = part_images.size();
                        part_images.push_back(images[i]);
                    }
                std::vector<LinkPair> part_pairs(sh.pair_ids.size());
                for (size_t k = 0; k < sh.pair_ids.size(); k++)
                    part_pairs[k] = LinkPair{local[pairs[sh.pair_ids[k]].image_1], local[pairs[sh.pair_ids[k]].image_2]};
                LinkOptions opt = options;
                opt.threads = std::max(1, all_threads / n_devices);
                opt.packed_out = nullptr, opt.packed_offsets = opt.packed_counts = nullptr; // per-part lists are not packed
                std::vector<opencalibration::camera_relations> part =
                    link_pairs(part_images, part_pairs, opt, &part_stats[(size_t)d]);
                for (size_t k = 0; k < sh.pair_ids.size(); k++)
                    relations[sh.pair_ids[k]] = std::move(part[k]); // serial order, link_stage.cpp:119-131
            }
            catch (...)
            {
                errors[(size_t)d] = std::current_exception();
            }
        });
    for (std::thread &w : workers)
        w.join();
    ocb_set_device(caller_device);
    for (const std::exception_ptr &e : errors)
        if (e)
            std::rethrow_exception(e);
    if (stats)
    {
        LinkStats st;
        for (const LinkStats &ps : part_stats)
        {
            st.seconds_subsample_upload += ps.seconds_subsample_upload, st.seconds_match_gpu += ps.seconds_match_gpu;
            st.seconds_tail += ps.seconds_tail, st.seconds_setup += ps.seconds_setup;
            st.seconds_release += ps.seconds_release;
            st.comparisons += ps.comparisons, st.matches += ps.matches, st.ransac_inliers += ps.ransac_inliers;
        }
        st.seconds_total = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        *stats = st;
    }
    return relations;
}
} // namespace ocb_host
