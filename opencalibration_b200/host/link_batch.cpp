// Batched LinkStage runner -- see link_batch.hpp. Reference: src/pipeline/link_stage.cpp:41-117.
#include "link_batch.hpp"

#include <cstdio>
#include <cstdlib>
#include "models_detail.hpp"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstring>
#include <condition_variable>
#include <deque>
#include <functional>
#include <future>
#include <limits>
#include <map>
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

// page-locked result buffers (one per slot of a link_pairs call), recycled through a small pool
struct ResultBuffers
{
    std::vector<void *> p;
    size_t cap = 0;
};
std::mutex g_pool_mu;
std::vector<ResultBuffers> g_pool;

void free_result_buffers(ResultBuffers &b)
{
    for (void *q : b.p)
        ocb_host_free(q);
    b.p.clear();
    b.cap = 0;
}

ResultBuffers take_result_buffers(size_t count, size_t bytes)
{
    ResultBuffers b;
    {
        std::lock_guard<std::mutex> lk(g_pool_mu);
        if (!g_pool.empty())
        {
            b = g_pool.back();
            g_pool.pop_back();
        }
    }
    if (b.cap < bytes)
        free_result_buffers(b);
    if (b.p.empty())
        b.cap = bytes + bytes / 4;
    while (b.p.size() < count)
    {
        void *q = ocb_host_alloc(b.cap);
        if (!q)
        {
            free_result_buffers(b);
            return b;
        }
        b.p.push_back(q);
    }
    while (b.p.size() > count)
    {
        ocb_host_free(b.p.back());
        b.p.pop_back();
    }
    return b;
}

void give_back_result_buffers(ResultBuffers &b)
{
    if (b.p.empty())
        return;
    std::lock_guard<std::mutex> lk(g_pool_mu);
    g_pool.push_back(b);
    b.p.clear();
}

inline double as_distance(uint32_t d)
{
    // distance = count * (1.0 / DESCRIPTOR_BITS)  (match_features.cpp:79)
    return d == OCB_DIST_INF ? std::numeric_limits<double>::infinity() : d * (1.0 / feature_2d::DESCRIPTOR_BITS);
}

ocb_camera camera_of(const DifferentiableCameraModel<double> &m)
{
    ocb_camera c;
    std::memset(&c, 0, sizeof c);
    c.focal_length_pixels = m.focal_length_pixels;
    c.principal_point[0] = m.principle_point[0], c.principal_point[1] = m.principle_point[1];
    for (int i = 0; i < 3; i++)
        c.radial_distortion[i] = m.radial_distortion[i];
    c.tangential_distortion[0] = m.tangential_distortion[0], c.tangential_distortion[1] = m.tangential_distortion[1];
    c.projection_planar = m.projection_type == ProjectionType::PLANAR ? 1 : 0;
    return c;
}

// the end of one LinkStage closure, after ransac (link_stage.cpp:95-108)
void finish_pair(const ocb_host::LinkImage &img, const ocb_host::LinkImage &near_image, camera_relations &relations,
                 std::vector<feature_match> &&coarse_matches, const std::vector<correspondence> &coarse_correspondences,
                 homography_model &h, const std::vector<bool> &coarse_inliers, size_t *n_inliers)
{
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

// Long-lived helper threads for link_pairs (the submission threads and the extra tail consumers), one pool per device.
// libocb keeps a per-thread context (two streams, device and page-locked staging areas, bound correspondences); a
// std::thread per call would build and tear that context down every time, which costs tens of milliseconds, and a
// thread that alternates between devices would rebuild it on every switch. The threads are created on demand, parked
// between calls and never joined (the pools are leaked on purpose: they must not run destructors after the CUDA
// runtime has shut down).
class HelperThreads
{
  public:
    // Role k of every call runs on the SAME thread (the subsample / upload thread, submission thread 0, submission
    // thread 1, tail consumer 1, ...): that thread's libocb context -- streams, staging areas sized for exactly this
    // role -- is warm from the previous call. (A shared queue let any parked thread take any role, and every thread
    // then had to grow its own page-locked staging area once per role: tens of milliseconds on a box where such an
    // allocation is mapped into eight GPUs, right at the start of a call, when nothing else can run.)
    std::future<void> run(size_t role, std::function<void()> fn)
    {
        std::packaged_task<void()> task(std::move(fn));
        std::future<void> done = task.get_future();
        Worker *w;
        {
            std::lock_guard<std::mutex> lk(mu_);
            while (workers_.size() <= role)
                workers_.push_back(nullptr);
            if (!workers_[role])
            {
                workers_[role] = new Worker;
                std::thread(&HelperThreads::loop, workers_[role]).detach();
            }
            w = workers_[role];
        }
        {
            std::lock_guard<std::mutex> lk(w->mu);
            w->queue.push_back(std::move(task));
        }
        w->cv.notify_one();
        return done;
    }
    static HelperThreads &instance(int device)
    {
        static std::mutex mu;
        static std::map<int, HelperThreads *> *pools = new std::map<int, HelperThreads *>;
        std::lock_guard<std::mutex> lk(mu);
        HelperThreads *&p = (*pools)[device];
        if (!p)
            p = new HelperThreads;
        return *p;
    }

  private:
    struct Worker
    {
        std::mutex mu;
        std::condition_variable cv;
        std::deque<std::packaged_task<void()>> queue;
    };
    static void loop(Worker *w)
    {
        for (;;)
        {
            std::packaged_task<void()> task;
            {
                std::unique_lock<std::mutex> lk(w->mu);
                w->cv.wait(lk, [&] { return !w->queue.empty(); });
                task = std::move(w->queue.front());
                w->queue.pop_front();
            }
            task();
        }
    }
    std::mutex mu_;
    std::vector<Worker *> workers_;
};

// The helper threads of one link_pairs call work on that call's stack variables. Whatever way the call is left
// (an allocation failure between two stages included), this guard tells them to drain and waits for them first.
struct HelperGuard
{
    std::mutex &mu;
    std::condition_variable &cv;
    bool &stop;
    std::vector<std::future<void>> tasks;
    ~HelperGuard()
    {
        {
            std::lock_guard<std::mutex> lk(mu);
            stop = true;
        }
        cv.notify_all();
        for (std::future<void> &t : tasks)
            if (t.valid())
                t.wait();
    }
};
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
    {
        if (!im.features)
            throw std::invalid_argument("link_pairs: image without features");
        // link_stage.cpp:63-65 reads features[0 .. num_sparse_features) unchecked; a count from a corrupt checkpoint
        // must not become an out-of-bounds read of descriptor memory
        if (im.num_sparse_features > im.features->size())
            throw std::invalid_argument("link_pairs: num_sparse_features exceeds the number of features");
    }
    // the helper threads below work on the device the CALLER selected (ocb_set_device), not on the process default
    const int device = ocb_current_device();
    const bool device_tail = options.device_tail;
    bool device_sort = device_tail && options.device_sort;
    if (const char *e = std::getenv("OCB_LINK_DEVICE_SORT")) // experiments / tests
        device_sort = device_tail && e[0] == '1';

    // ---- submissions. The first ones are small and grow geometrically up to pairs_per_submission: the GPU starts on
    // the first pairs as soon as THEIR images are resident, and the first results reach the tail workers early. The
    // last ones shrink the same way: what is left to do after the GPU has finished is the tail of a small submission.
    // In between every submission is large enough to fill the SMs for tens of milliseconds.
    int tail_workers = options.tail_workers;
    if (const char *e = std::getenv("OCB_LINK_TAIL_WORKERS")) // experiments
        tail_workers = std::max(1, std::atoi(e));
    const int workers = options.run_ransac ? std::max(1, std::min(tail_workers, threads)) : 1;
    const int team = std::max(1, threads / workers);
    // With RANSAC a submission's tail is a chain of lock-step rounds on ONE tail worker's team; when that team is a
    // single thread (few cores per GPU) a large submission would keep its worker busy long after the GPU has moved on,
    // so the submissions are kept small enough for the workers to share the load evenly.
    size_t per = std::max<size_t>(1, options.pairs_per_submission);
    if (options.run_ransac && team < 2)
        per = std::min<size_t>(per, 64);
    std::vector<std::pair<size_t, size_t>> chunk; // [begin, end) into `pairs`
    {
        std::vector<size_t> ramp; // 32, 64, 128, ... below `per`
        for (size_t size = std::max<size_t>(1, options.first_submission); size < per; size *= 2)
            ramp.push_back(size);
        size_t ramp_total = 0;
        for (size_t r : ramp)
            ramp_total += r;
        std::vector<size_t> sizes;
        if (n_pairs > 2 * ramp_total + per)
        {
            sizes = ramp;
            const size_t middle = n_pairs - 2 * ramp_total, k = (middle + per - 1) / per;
            for (size_t i = 0; i < k; i++)
                sizes.push_back(middle / k + (i < middle % k ? 1 : 0)); // equal middle submissions
            sizes.insert(sizes.end(), ramp.rbegin(), ramp.rend());
        }
        else // short pair list: ramp up only
            for (size_t left = n_pairs, size = ramp.empty() ? per : ramp.front(); left;)
            {
                const size_t take = std::min(left, size);
                sizes.push_back(take);
                left -= take;
                size = std::min(per, size * 2);
            }
        size_t begin = 0;
        for (size_t sz : sizes)
        {
            chunk.emplace_back(begin, begin + sz);
            begin += sz;
        }
    }
    const size_t n_chunks = chunk.size();

    // ---- per image, once: the subsample every closure of that image would compute (link_stage.cpp:63-65,80-81) and
    // the upload of those rows (+ keypoints and camera model for the device-side rays). Only images that occur in a
    // pair are touched. The images are prepared in the order in which the submissions need them, by a helper thread
    // that runs ahead of the matching.
    std::vector<size_t> order;                       // used images, by first use
    std::vector<size_t> position(n_img, ~(size_t)0); // image -> index in `order`
    std::vector<size_t> chunk_needs(n_chunks, 0);    // images of `order` that must be resident before chunk c starts
    for (size_t c = 0; c < n_chunks; c++)
    {
        for (size_t p = chunk[c].first; p < chunk[c].second; p++)
            for (size_t i : {pairs[p].image_1, pairs[p].image_2})
                if (position[i] == ~(size_t)0)
                {
                    position[i] = order.size();
                    order.push_back(i);
                }
        chunk_needs[c] = order.size();
    }
    std::vector<std::vector<size_t>> indices(n_img);
    const uint64_t id_base = g_next_set_id.fetch_add(n_img + 1);
    std::string error;
    std::vector<char> registered(n_img, 0);
    std::mutex mu;
    std::condition_variable cv;
    bool stop = false;   // set on the first error: everybody drains
    size_t prepared = 0; // images of `order` that are subsampled and resident
    double prepare_seconds = 0;
    // timeline of every submission, seconds since the call began (OCB_LINK_TRACE=2 prints it)
    struct ChunkTimes
    {
        double ready = 0, launch = 0, matched = 0, tail_begin = 0, tail_end = 0;
    };
    std::vector<ChunkTimes> timeline(n_chunks);
    auto prepare = [&]() {
        ocb_set_device(device);
        const size_t batch = (size_t)std::max(32, 4 * threads);
        size_t begin = 0;
        for (size_t c = 0; c < n_chunks; c++)
        {
            // everything chunk c still needs, in batches (a batch is one device allocation and one pipelined upload)
            while (begin < chunk_needs[c])
            {
                {
                    std::lock_guard<std::mutex> lk(mu);
                    if (stop)
                        return;
                }
                const auto t0 = clock_type::now();
                const size_t end = std::min(chunk_needs[c], begin + batch);
                std::string local_error;
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
                for (size_t u = begin; u < end; u++)
                {
                    const size_t i = order[u];
                    try
                    {
                        indices[i] = spatially_subsample_feature_indices(
                            *images[i].features, options.coarse_spacing_pixels, images[i].num_sparse_features);
                    }
                    catch (const std::exception &e)
                    {
#pragma omp critical(ocb_link_error)
                        local_error = e.what();
                    }
                }
                // upload: one batched registration, rows and keypoints taken straight from the feature vectors
                // through the subsample indices
                if (local_error.empty())
                {
                    std::vector<ocb_image_source> src;
                    for (size_t u = begin; u < end; u++)
                    {
                        const size_t i = order[u];
                        const std::vector<feature_2d> &f = *images[i].features;
                        ocb_image_source s;
                        std::memset(&s, 0, sizeof s);
                        s.set_id = id_base + i;
                        s.rows = f.empty() ? nullptr : static_cast<const void *>(&f[0].descriptor);
                        s.stride = sizeof(feature_2d);
                        s.idx = indices[i].data();
                        s.n = indices[i].size();
                        if (device_tail && options.run_ransac)
                        {
                            s.xy = f.empty() ? nullptr : static_cast<const void *>(&f[0].location);
                            s.xy_stride = sizeof(feature_2d);
                            s.camera = camera_of(images[i].model);
                        }
                        src.push_back(s);
                    }
                    if (ocb_register_images_batch(src.data(), src.size()))
                        local_error = std::string("ocb_register_images_batch: ") + ocb_last_error();
                    else
                        for (const ocb_image_source &sset : src)
                            registered[sset.set_id - id_base] = 1;
                }
                std::lock_guard<std::mutex> lk(mu);
                prepare_seconds += since(t0);
                if (!local_error.empty())
                {
                    if (error.empty())
                        error = local_error;
                    stop = true;
                }
                else
                {
                    prepared = end;
                    for (size_t cc = 0; cc < n_chunks; cc++)
                        if (timeline[cc].ready == 0 && chunk_needs[cc] <= prepared)
                            timeline[cc].ready = since(t_begin);
                }
                cv.notify_all();
                if (stop)
                    return;
                begin = end;
            }
        }
    };
    auto release_sets = [&]() {
        for (size_t i = 0; i < n_img; i++)
            if (registered[i])
                ocb_unregister_descriptors(id_base + i);
    };
    LinkStats st;

    // ---- producers (parked helper threads, see HelperThreads) keep the GPU matching the next submissions (into
    // page-locked result buffers, one per slot) while `tail_workers` consumer threads, each with its own OpenMP team,
    // finish the submissions already matched. Several consumers are needed because a submission's RANSAC rounds are a
    // serial chain of GPU round trips: with one consumer the chain's latency, not the host cores, bounds the tail.
    size_t max_rows = 1; // upper bound: the subsample keeps at most the sparse features of an image
    for (size_t c = 0; c < n_chunks; c++)
    {
        size_t rows = 0;
        for (size_t p = chunk[c].first; p < chunk[c].second; p++)
        {
            const LinkImage &im = images[pairs[p].image_1];
            rows += im.num_sparse_features ? im.num_sparse_features : im.features->size();
        }
        max_rows = std::max(max_rows, rows);
    }
    // two submissions in flight (two producers on alternate chunks, each with its own stream): while one submission
    // drains its last CTAs, returns its records and the next problem table is built, the other one fills the SMs
    // (three with the device sort: a submission then ends with the sequential sort replay K7, about three milliseconds
    // during which its stream has little for the SMs; the submission threads sleep while they wait)
    size_t want_producers = device_sort ? 3 : 2;
    if (const char *e = std::getenv("OCB_LINK_PRODUCERS")) // experiments
        want_producers = (size_t)std::max(1, std::atoi(e));
    const size_t n_producers = std::min<size_t>(want_producers, std::max<size_t>(n_chunks, 1));
    const size_t n_slots = (size_t)workers + n_producers;
    struct Slot
    {
        void *records = nullptr;       // device_tail: ocb_match survivors (dense, sorted); else ocb_top2 per query row
        std::vector<uint32_t> quality_order; // device_tail + RANSAC: PROSAC order of every pair's sorted matches
        std::vector<uint64_t> offsets; // device_tail: [pairs + 1] survivor offsets; else [pairs] row offsets
        double gpu_seconds = 0;
        size_t chunk = 0; // which submission the records belong to
        size_t next = 0;  // the submission this slot takes next: k, k + n_slots, ... strictly in that order, whichever
                          // producer owns them (two producers on alternate chunks are not ordered against each other)
        bool full = false;
        bool busy = false; // claimed by a producer that is still matching into it
    };
    std::vector<Slot> slot(n_slots);
    for (size_t k = 0; k < n_slots; k++)
        slot[k].next = k;
    // page-locked result buffers are expensive to create (the driver maps them into every visible GPU), so they are
    // kept between calls and only grow
    const size_t record_bytes = device_tail ? sizeof(ocb_match) : sizeof(ocb_top2);
    ResultBuffers buffers = take_result_buffers(n_slots, max_rows * record_bytes);
    if (buffers.p.size() != n_slots)
    {
        give_back_result_buffers(buffers);
        throw std::runtime_error(std::string("ocb_host_alloc: ") + ocb_last_error());
    }
    for (size_t k = 0; k < n_slots; k++)
        slot[k].records = buffers.p[k];
    st.seconds_setup = since(t_begin);
    auto produce = [&](size_t first) {
        ocb_set_device(device);
        for (size_t c = first; c < n_chunks; c += n_producers)
        {
            Slot &sl = slot[c % n_slots];
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] {
                    return (!sl.full && !sl.busy && sl.next == c && prepared >= chunk_needs[c]) || stop;
                });
                if (stop)
                    return;
                sl.busy = true;
            }
            const size_t begin = chunk[c].first, end = chunk[c].second;
            std::vector<ocb_pair> sub(end - begin);
            sl.offsets.assign(sub.size() + 1, 0);
            uint64_t total = 0;
            for (size_t p = begin; p < end; p++)
            {
                sub[p - begin] = ocb_pair{id_base + pairs[p].image_1, id_base + pairs[p].image_2};
                sl.offsets[p - begin] = total;
                total += indices[pairs[p].image_1].size();
            }
            const auto t0 = clock_type::now();
            timeline[c].launch = since(t_begin);
            int rc;
            if (device_tail)
            {
                // K1 + ratio test + compaction on the device: only the survivors come back; with device_sort already in
                // the order of match_features.cpp:100-101 and with their PROSAC order when RANSAC follows
                if (device_sort)
                {
                    if (options.run_ransac)
                        sl.quality_order.resize(max_rows);
                    rc = ocb_match_pairs_sorted(sub.data(), sub.size(), static_cast<ocb_match *>(sl.records), max_rows,
                                                sl.offsets.data(), options.run_ransac ? sl.quality_order.data() : nullptr);
                }
                else
                    rc = ocb_match_pairs_ratio(sub.data(), sub.size(), static_cast<ocb_match *>(sl.records), max_rows,
                                               sl.offsets.data());
            }
            else
                rc = ocb_match_pairs(sub.data(), sub.size(), static_cast<ocb_top2 *>(sl.records), sl.offsets.data());
            sl.gpu_seconds = since(t0);
            timeline[c].matched = since(t_begin);
            std::lock_guard<std::mutex> lk(mu);
            if (rc)
            {
                if (error.empty())
                    error = std::string(device_tail ? "ocb_match_pairs_sorted: " : "ocb_match_pairs: ") + ocb_last_error();
                stop = true;
            }
            sl.chunk = c;
            sl.full = true;
            sl.busy = false;
            cv.notify_all();
            if (rc)
                return;
        }
    };

    std::vector<camera_relations> relations(n_pairs);
    std::atomic<size_t> next_chunk{0}, total_matches{0}, total_inliers{0};
    std::atomic<uint64_t> packed_cursor{0};
    double tail_seconds = 0, gpu_seconds = 0;
    double phase_seconds[3] = {0, 0, 0}; // consumer wall time in (a) ratio test / sort / rays, (b) RANSAC, (c) finish
    auto fail = [&](const std::string &what) {
        std::lock_guard<std::mutex> lk(mu);
        if (error.empty())
            error = what;
        stop = true;
        cv.notify_all();
    };
    auto consume = [&]() {
        ocb_set_device(device);
        // a tail worker that waits for a RANSAC round gives its core to the other workers instead of spinning on it
        // (the caller's own setting is restored when the call returns)
        ocb_set_thread_blocking_sync(threads < omp_get_num_procs() || options.tail_workers > 1 ? 1 : 0);
        for (;;)
        {
            const size_t c = next_chunk.fetch_add(1);
            if (c >= n_chunks)
                return;
            Slot &sl = slot[c % n_slots];
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return (sl.full && sl.chunk == c) || stop; });
                if (stop)
                    return;
            }
            const size_t begin = chunk[c].first, end = chunk[c].second, cn = end - begin;
            const auto t0 = clock_type::now();
            timeline[c].tail_begin = since(t_begin);
            std::string local_error;
            size_t n_matches = 0, n_inl = 0;
            // (a) per pair: [ratio test +] the reference's std::sort -> match list (link_stage.cpp:83-85);
            //     rays (:87-88) on the host, or left to the device (K6) when the RANSAC runs are bound
            std::vector<std::vector<feature_match>> coarse_matches(cn);
            std::vector<std::vector<correspondence>> coarse_correspondences(cn);
            std::vector<std::vector<ocb_match>> sorted(device_tail && options.run_ransac ? cn : 0);
            std::vector<std::vector<uint32_t>> quality_order(sorted.size());
#pragma omp parallel for schedule(dynamic, 1) num_threads(team) reduction(+ : n_matches)
            for (size_t p = begin; p < end; p++)
            {
                try
                {
                    const size_t k = p - begin;
                    const LinkImage &img = images[pairs[p].image_1], &near_image = images[pairs[p].image_2];
                    const std::vector<size_t> &idx1 = indices[pairs[p].image_1], &idx2 = indices[pairs[p].image_2];
                    if (device_tail)
                    {
                        // Survivors are emitted in query order (match_features.cpp:71-97) and then put through
                        // libstdc++'s std::sort on the distance (:100-101) -- here, or already on the device (K7).
                        // std::sort's sequence of comparisons and moves depends only on the comparator's answers, and
                        // a.d > b.d <=> a.d * (1.0 / 486) > b.d * (1.0 / 486) for these integers, so sorting the
                        // 12-byte records yields the permutation the reference's sort produces on its 24-byte ones.
                        const ocb_match *first = static_cast<const ocb_match *>(sl.records) + sl.offsets[k];
                        std::vector<ocb_match> recs(first, first + (sl.offsets[k + 1] - sl.offsets[k]));
                        if (!device_sort)
                            std::sort(recs.begin(), recs.end(),
                                      [](const ocb_match &f1, const ocb_match &f2) -> bool { return f1.best_d > f2.best_d; });
                        std::vector<feature_match> &out = coarse_matches[k];
                        out.resize(recs.size());
                        for (size_t i = 0; i < recs.size(); i++)
                            out[i] = feature_match{idx1[recs[i].query_k], idx2[recs[i].best_k], as_distance(recs[i].best_d)};
                        if (options.run_ransac)
                        {
                            // the rays are computed on the device when the runs are bound; the driver only needs the
                            // qualities beforehand (PROSAC order, ransac.cpp:72-90)
                            std::vector<correspondence> &corr = coarse_correspondences[k];
                            corr.resize(recs.size());
                            for (size_t i = 0; i < recs.size(); i++)
                                corr[i].quality = out[i].distance;
                            if (recs.size() < homography_model::MINIMUM_POINTS) // never bound (ransac.cpp:64-70)
                                corr = distort_keypoints(*img.features, *near_image.features, out, img.model,
                                                         near_image.model);
                            if (device_sort)
                                quality_order[k].assign(sl.quality_order.begin() + sl.offsets[k],
                                                        sl.quality_order.begin() + sl.offsets[k + 1]);
                            sorted[k] = std::move(recs);
                        }
                    }
                    else
                    {
                        coarse_matches[k] = detail::matches_from_top2(
                            idx1, idx2, static_cast<const ocb_top2 *>(sl.records) + sl.offsets[k], nullptr, nullptr);
                        if (options.run_ransac)
                            coarse_correspondences[k] = distort_keypoints(*img.features, *near_image.features,
                                                                          coarse_matches[k], img.model, near_image.model);
                    }
                    n_matches += coarse_matches[k].size();
                    if (!options.run_ransac)
                        relations[p].matches = std::move(coarse_matches[k]);
                }
                catch (const std::exception &e)
                {
#pragma omp critical(ocb_link_error)
                    local_error = e.what();
                }
            }
            const double gpu_s = sl.gpu_seconds;
            const double t_a = since(t0);
            double t_b = t_a;
            {
                // the records have been consumed: the slot can take the next submission
                std::lock_guard<std::mutex> lk(mu);
                sl.full = false;
                sl.next = c + n_slots;
                cv.notify_all();
            }
            if (options.run_ransac && local_error.empty())
            {
                // (b) the RANSAC runs of all pairs of the submission advance in lock step: one launch per round
                std::vector<homography_model> models(cn);
                std::vector<std::vector<bool>> coarse_inliers(cn);
                std::vector<RansacJob<homography_model>> jobs(cn);
                for (size_t k = 0; k < cn; k++)
                {
                    jobs[k].matches = &coarse_correspondences[k], jobs[k].model = &models[k],
                    jobs[k].inliers = &coarse_inliers[k];
                    if (device_tail && !quality_order[k].empty())
                        jobs[k].quality_order = quality_order[k].data();
                }
                // device tail: bind the runs straight from the sorted match lists -- K6 computes both rays of every
                // match (distort_keypoints, link_stage.cpp:87-88) into the bound rows and the rows come back for
                // decompose(); 12 bytes per match go up instead of 56
                const CorrBinder bind_from_matches = [&](const std::vector<ocb_corr_set> &sets) {
                    std::vector<ocb_match_set> ms(sets.size());
                    for (size_t k = 0; k < sets.size(); k++)
                    {
                        std::memset(&ms[k], 0, sizeof ms[k]);
                        if (sets[k].n == 0)
                            continue;
                        ms[k].set_1 = id_base + pairs[begin + k].image_1, ms[k].set_2 = id_base + pairs[begin + k].image_2;
                        ms[k].matches = sorted[k].data(), ms[k].n = sorted[k].size(), ms[k].order = sets[k].order;
                        ms[k].corr_out = reinterpret_cast<double *>(coarse_correspondences[k].data());
                    }
                    detail::gpu_check(ocb_corr_bind_batch_matches(ms.data(), ms.size()), "ocb_corr_bind_batch_matches");
                };
                try
                {
                    ransac_batch(jobs, team, device_tail ? &bind_from_matches : nullptr);
                }
                catch (const std::exception &e)
                {
                    local_error = e.what();
                }
                t_b = since(t0);
                // (c) decomposition + inlier assembly (link_stage.cpp:95-108)
                if (local_error.empty())
                {
#pragma omp parallel for schedule(dynamic, 1) num_threads(team) reduction(+ : n_inl)
                    for (size_t p = begin; p < end; p++)
                    {
                        try
                        {
                            size_t inl = 0;
                            finish_pair(images[pairs[p].image_1], images[pairs[p].image_2], relations[p],
                                        std::move(coarse_matches[p - begin]), coarse_correspondences[p - begin],
                                        models[p - begin], coarse_inliers[p - begin], &inl);
                            n_inl += inl;
                        }
                        catch (const std::exception &e)
                        {
#pragma omp critical(ocb_link_error)
                            local_error = e.what();
                        }
                    }
                }
            }
            // (d) the flat copy of the final match lists, if the caller gathers them
            if (options.packed_out && local_error.empty())
            {
                std::vector<uint64_t> at(cn + 1, 0);
                for (size_t k = 0; k < cn; k++)
                    at[k + 1] = at[k] + relations[begin + k].matches.size();
                const uint64_t base = packed_cursor.fetch_add(at[cn]);
                if (base + at[cn] > options.packed_capacity)
                    local_error = "link_pairs: packed_capacity is smaller than the number of matches";
                else
                {
#pragma omp parallel for schedule(dynamic, 4) num_threads(team)
                    for (size_t k = 0; k < cn; k++)
                    {
                        uint32_t *out = options.packed_out + 3 * (base + at[k]);
                        for (const feature_match &m : relations[begin + k].matches)
                        {
                            out[0] = (uint32_t)m.feature_index_1, out[1] = (uint32_t)m.feature_index_2;
                            out[2] = (uint32_t)std::lround(m.distance * feature_2d::DESCRIPTOR_BITS);
                            out += 3;
                        }
                        if (options.packed_offsets)
                            options.packed_offsets[begin + k] = base + at[k];
                        if (options.packed_counts)
                            options.packed_counts[begin + k] = at[k + 1] - at[k];
                    }
                }
            }
            total_matches += n_matches;
            total_inliers += n_inl;
            timeline[c].tail_end = since(t_begin);
            {
                std::lock_guard<std::mutex> lk(mu);
                const double t_c = since(t0);
                tail_seconds += t_c;
                gpu_seconds += gpu_s;
                phase_seconds[0] += t_a, phase_seconds[1] += t_b - t_a, phase_seconds[2] += t_c - t_b;
            }
            if (!local_error.empty())
            {
                fail(local_error);
                return;
            }
        }
    };
    {
        // Everything the helper tasks touch is declared above; the guard is the innermost object, so whichever way
        // this scope is left it stops and joins the tasks before any of that state goes away.
        HelperGuard helpers{mu, cv, stop, {}};
        HelperThreads &pool = HelperThreads::instance(device);
        helpers.tasks.reserve(2 + n_producers + (size_t)workers);
        helpers.tasks.push_back(pool.run(0, prepare));
        for (size_t k = 0; k < n_producers; k++)
            helpers.tasks.push_back(pool.run(1 + k, [&produce, k]() { produce(k); }));
        const size_t first_consumer = helpers.tasks.size();
        for (int w = 1; w < workers; w++)
            helpers.tasks.push_back(pool.run(8 + (size_t)w, consume)); // roles 1..8: submission threads
        consume();
        ocb_set_thread_blocking_sync(0);
        for (size_t k = first_consumer; k < helpers.tasks.size(); k++)
            helpers.tasks[k].wait(); // every chunk is consumed (or an error stopped the run); the guard releases the rest
    }
    st.seconds_subsample_upload = prepare_seconds; // overlapped with the matching after the first submission
    st.seconds_match_gpu = gpu_seconds; // summed over submissions, two of which are in flight at a time
    st.seconds_tail = tail_seconds;
    const auto t_release = clock_type::now();
    give_back_result_buffers(buffers);
    release_sets();
    st.seconds_release = since(t_release);
    if (!error.empty())
        throw std::runtime_error(error);
    for (const LinkPair &p : pairs)
        st.comparisons += indices[p.image_1].size() * indices[p.image_2].size();
    st.matches = total_matches.load();
    st.ransac_inliers = total_inliers.load();
    st.seconds_total = since(t_begin);
    if (std::getenv("OCB_LINK_TRACE"))
        std::fprintf(stderr,
                     "link_pairs: %zu pairs, %zu chunks, %d consumers x %d threads: upload %.3f s, match (GPU, summed) %.3f s, "
                     "tail %.3f s = sort/rays %.3f + ransac %.3f + finish %.3f, total %.3f s\n",
                     n_pairs, n_chunks, workers, team, st.seconds_subsample_upload, gpu_seconds, tail_seconds,
                     phase_seconds[0], phase_seconds[1], phase_seconds[2], st.seconds_total);
    if (const char *tr = std::getenv("OCB_LINK_TRACE"))
        if (tr[0] == '2')
            for (size_t c = 0; c < n_chunks; c++)
                std::fprintf(stderr, "  chunk %2zu: %4zu pairs, images ready %.4f, launch %.4f, matched %.4f, tail %.4f .. %.4f\n", c,
                             chunk[c].second - chunk[c].first, timeline[c].ready, timeline[c].launch, timeline[c].matched,
                             timeline[c].tail_begin, timeline[c].tail_end);
    if (stats)
        *stats = st;
    return relations;
}
} // namespace ocb_host
