// Flat C entry points over the C++ mirror, for ctypes-driven tests and bench.py (opencalibration_b200/host.py).
// They build the reference-typed arguments (std::vector<feature_2d>, std::vector<correspondence>, model structs),
// call the mirror exactly as src/pipeline/link_stage.cpp:80-93 would, and flatten the results.
#include "guided_match.hpp"
#include "link_batch.hpp"
#include "partition.hpp"
#include "models_detail.hpp"

#include <algorithm>
#include <chrono>
#include <cstring>
#include <omp.h>
#include <stdexcept>
#include <string>

using namespace opencalibration;

namespace
{
thread_local std::string t_err;

std::vector<feature_2d> make_features(const double *xy, const float *strength, const uint64_t *desc, size_t n)
{
    std::vector<feature_2d> f(n);
    for (size_t i = 0; i < n; i++)
    {
        if (xy)
            f[i].location = Eigen::Vector2d(xy[2 * i], xy[2 * i + 1]);
        if (strength)
            f[i].strength = strength[i];
        if (desc)
            std::memcpy(static_cast<void *>(&f[i].descriptor), desc + 8 * i, 64);
    }
    return f;
}
std::vector<correspondence> make_corr(const double *corr, size_t n)
{
    std::vector<correspondence> c(n);
    if (n)
        std::memcpy(static_cast<void *>(c.data()), corr, n * sizeof(correspondence));
    return c;
}
std::vector<bool> make_flags(const uint8_t *f, size_t n)
{
    std::vector<bool> v(n);
    for (size_t i = 0; i < n; i++)
        v[i] = f[i] != 0;
    return v;
}
void load(homography_model &m, const double *M18)
{
    std::memcpy(m.homography.data(), M18, 72);
    std::memcpy(m.homography_inverse.data(), M18 + 9, 72);
}
void load(essential_matrix_model &m, const double *M18)
{
    std::memcpy(m.essential_matrix.data(), M18, 72);
}
void load(fundamental_matrix_model &m, const double *M18)
{
    std::memcpy(m.fundamental_matrix.data(), M18, 72);
}
template <typename M> void store(const M &m, double *M18)
{
    ocb_host::detail::pack_model(m, M18);
}

// run f on a freshly built model of the requested kind
template <typename F> auto with_model(int kind, const double *M18, double thr, F &&f)
{
    if (kind == OCB_MODEL_HOMOGRAPHY)
    {
        homography_model m;
        if (M18)
            load(m, M18);
        if (thr > 0)
            m.inlier_threshold = thr;
        return f(m);
    }
    if (kind == OCB_MODEL_ESSENTIAL)
    {
        essential_matrix_model m;
        if (M18)
            load(m, M18);
        if (thr > 0)
            m.inlier_threshold = thr;
        return f(m);
    }
    fundamental_matrix_model m;
    if (M18)
        load(m, M18);
    if (thr > 0)
        m.inlier_threshold = thr;
    return f(m);
}

template <typename F> int guarded(F &&f)
{
    try
    {
        f();
        return 0;
    }
    catch (const std::exception &e)
    {
        t_err = e.what();
        return -1;
    }
}
} // namespace

namespace
{
// ransac_batch over n_jobs correspondence sets stored back to back (offsets[j] .. offsets[j+1]) -> per job the
// returned score, the model (18 doubles), the inlier flags (same layout as corr) and (iterations, improvements)
template <typename Model> void run_batch(const double *corr, const size_t *offsets, size_t n_jobs, int threads,
                                                double *scores, double *M18, uint8_t *inl, size_t *stats2)
{
    std::vector<std::vector<correspondence>> c(n_jobs);
    std::vector<Model> models(n_jobs);
    std::vector<std::vector<bool>> flags(n_jobs);
    std::vector<ocb_host::RansacJob<Model>> jobs(n_jobs);
    for (size_t j = 0; j < n_jobs; j++)
    {
        c[j] = make_corr(corr + 7 * offsets[j], offsets[j + 1] - offsets[j]);
        jobs[j].matches = &c[j], jobs[j].model = &models[j], jobs[j].inliers = &flags[j];
    }
    ocb_host::ransac_batch(jobs, threads);
    for (size_t j = 0; j < n_jobs; j++)
    {
        scores[j] = jobs[j].result;
        store(models[j], M18 + 18 * j);
        for (size_t k = 0; k < flags[j].size(); k++)
            inl[offsets[j] + k] = flags[j][k];
        stats2[2 * j] = jobs[j].stats.iterations, stats2[2 * j + 1] = jobs[j].stats.improvements;
    }
}
} // namespace

extern "C"
{
    const char *ocbh_last_error() { return t_err.c_str(); }
    size_t ocbh_sizeof_feature_2d() { return sizeof(feature_2d); }
    size_t ocbh_offsetof_descriptor() { return offsetof(feature_2d, descriptor); }
    size_t ocbh_sizeof_feature_match() { return sizeof(feature_match); }
    size_t ocbh_sizeof_correspondence() { return sizeof(correspondence); }
    size_t ocbh_sizeof_feature_match_denormalized() { return sizeof(feature_match_denormalized); }

    // ---- src/match ------------------------------------------------------------------------------------
    // n_out receives the number of matches; out_* hold up to n1 entries; mutual nullable (cross-check flags)
    int ocbh_match_features_subset(const uint64_t *desc1, size_t nf1, const uint64_t *desc2, size_t nf2,
                                   const size_t *idx1, size_t n1, const size_t *idx2, size_t n2, size_t *out_i1,
                                   size_t *out_i2, double *out_dist, uint8_t *mutual, size_t *n_out)
    {
        return guarded([&] {
            const std::vector<feature_2d> f1 = make_features(nullptr, nullptr, desc1, nf1);
            const std::vector<feature_2d> f2 = make_features(nullptr, nullptr, desc2, nf2);
            const std::vector<size_t> i1(idx1, idx1 + n1), i2(idx2, idx2 + n2);
            std::vector<bool> mut;
            const std::vector<feature_match> r = mutual ? ocb_host::match_features_subset_cross_checked(f1, f2, i1, i2, mut)
                                                        : match_features_subset(f1, f2, i1, i2);
            for (size_t i = 0; i < r.size(); i++)
            {
                out_i1[i] = r[i].feature_index_1;
                out_i2[i] = r[i].feature_index_2;
                out_dist[i] = r[i].distance;
                if (mutual)
                    mutual[i] = mut[i];
            }
            *n_out = r.size();
        });
    }

    // Persistent feature sets: the caller-side std::vector<feature_2d> the pipeline already holds in memory
    // (image::features). Matching two handles is exactly the reference call on existing vectors.
    void *ocbh_features_create(const double *xy, const float *strength, const uint64_t *desc, size_t n)
    {
        return new std::vector<feature_2d>(make_features(xy, strength, desc, n));
    }
    void ocbh_features_destroy(void *h) { delete static_cast<std::vector<feature_2d> *>(h); }
    int ocbh_match_handles(const void *h1, const void *h2, const size_t *idx1, size_t n1, const size_t *idx2,
                           size_t n2, size_t *out_i1, size_t *out_i2, double *out_dist, uint8_t *mutual, size_t *n_out)
    {
        return guarded([&] {
            const auto &f1 = *static_cast<const std::vector<feature_2d> *>(h1);
            const auto &f2 = *static_cast<const std::vector<feature_2d> *>(h2);
            const std::vector<size_t> i1(idx1, idx1 + n1), i2(idx2, idx2 + n2);
            std::vector<bool> mut;
            const std::vector<feature_match> r = mutual ? ocb_host::match_features_subset_cross_checked(f1, f2, i1, i2, mut)
                                                        : match_features_subset(f1, f2, i1, i2);
            for (size_t i = 0; i < r.size(); i++)
            {
                out_i1[i] = r[i].feature_index_1;
                out_i2[i] = r[i].feature_index_2;
                out_dist[i] = r[i].distance;
                if (mutual)
                    mutual[i] = mut[i];
            }
            *n_out = r.size();
        });
    }

    // ---- src/dense guided matcher (guided_match.hpp). Lists in CSR form; outputs hold up to n_lists entries.
    int ocbh_match_guided(const uint64_t *desc1, size_t nf1, const uint64_t *desc2, size_t nf2,
                          const size_t *query_feature, const size_t *begin, const size_t *nearby, size_t n_lists,
                          size_t *out_list, size_t *out_query, size_t *out_candidate, double *out_best,
                          double *out_second, size_t *n_out)
    {
        return guarded([&] {
            const std::vector<feature_2d> f1 = make_features(nullptr, nullptr, desc1, nf1);
            const std::vector<feature_2d> f2 = make_features(nullptr, nullptr, desc2, nf2);
            ocb_host::GuidedLists lists;
            lists.query_feature.assign(query_feature, query_feature + n_lists);
            lists.begin.assign(begin, begin + n_lists + 1);
            lists.nearby.assign(nearby, nearby + begin[n_lists]);
            const std::vector<ocb_host::GuidedMatch> r = ocb_host::match_features_guided(f1, f2, lists);
            for (size_t i = 0; i < r.size(); i++)
            {
                out_list[i] = r[i].list;
                out_query[i] = r[i].query_feature;
                out_candidate[i] = r[i].candidate_feature;
                out_best[i] = r[i].best_distance;
                out_second[i] = r[i].second_distance;
            }
            *n_out = r.size();
        });
    }

    size_t ocbh_subsample(const double *xy, const float *strength, size_t n, double spacing, size_t count,
                          size_t *out_idx)
    {
        const std::vector<feature_2d> f = make_features(xy, strength, nullptr, n);
        const std::vector<size_t> r = spatially_subsample_feature_indices(f, spacing, count);
        std::memcpy(out_idx, r.data(), r.size() * sizeof(size_t));
        return r.size();
    }

    // ---- src/model_inliers ------------------------------------------------------------------------------
    int ocbh_ransac(int kind, const double *corr, size_t n, double *M18, uint8_t *inliers, double *score,
                    size_t *stats6)
    {
        return guarded([&] {
            const std::vector<correspondence> c = make_corr(corr, n);
            std::vector<bool> inl;
            *score = with_model(kind, nullptr, 0, [&](auto &m) {
                const double s = ransac(c, m, inl);
                store(m, M18);
                return s;
            });
            for (size_t i = 0; i < inl.size(); i++)
                inliers[i] = inl[i];
            if (stats6)
            {
                const ocb_host::RansacStats st = ocb_host::last_ransac_stats();
                stats6[0] = st.iterations, stats6[1] = st.improvements, stats6[2] = st.rejected;
                stats6[3] = st.degenerate, stats6[4] = st.scored, stats6[5] = st.gpu_calls;
            }
        });
    }

    void ocbh_set_ransac_device_fit(int on) { ocb_host::set_ransac_device_fit(on != 0); }
    int ocbh_ransac_device_fit() { return ocb_host::ransac_device_fit() ? 1 : 0; }

    int ocbh_evaluate(int kind, const double *M18, double thr, const double *corr, size_t n, uint8_t *inliers,
                      double *score)
    {
        return guarded([&] {
            const std::vector<correspondence> c = make_corr(corr, n);
            std::vector<bool> inl;
            *score = with_model(kind, M18, thr, [&](auto &m) { return m.evaluate(c, inl); });
            for (size_t i = 0; i < inl.size(); i++)
                inliers[i] = inl[i];
        });
    }

    double ocbh_error(int kind, const double *M18, const double *corr7)
    {
        correspondence c;
        std::memcpy(static_cast<void *>(&c), corr7, sizeof c);
        return with_model(kind, M18, 0, [&](auto &m) { return m.error(c); });
    }

    void ocbh_fit(int kind, const double *corr, size_t n, const size_t *sample, double *M18)
    {
        const std::vector<correspondence> c = make_corr(corr, n);
        with_model(kind, nullptr, 0, [&](auto &m) {
            using M = std::decay_t<decltype(m)>;
            std::array<size_t, M::MINIMUM_POINTS> s;
            for (size_t i = 0; i < M::MINIMUM_POINTS; i++)
                s[i] = sample[i];
            m.fit(c, s);
            store(m, M18);
            return 0;
        });
    }

    void ocbh_fit_inliers(int kind, double *M18, const double *corr, size_t n, const uint8_t *inliers)
    {
        const std::vector<correspondence> c = make_corr(corr, n);
        const std::vector<bool> inl = make_flags(inliers, n);
        with_model(kind, M18, 0, [&](auto &m) {
            m.fitInliers(c, inl);
            store(m, M18);
            return 0;
        });
    }

    int ocbh_check_sample_degeneracy_h(const double *corr, size_t n, const size_t *sample)
    {
        const std::vector<correspondence> c = make_corr(corr, n);
        return homography_model::checkSampleDegeneracy(c, {sample[0], sample[1], sample[2], sample[3]}) ? 1 : 0;
    }

    int ocbh_check_degeneracy_f(double *M18, double thr, const double *corr, size_t n, uint8_t *inliers)
    {
        return guarded([&] {
            const std::vector<correspondence> c = make_corr(corr, n);
            std::vector<bool> inl = make_flags(inliers, n);
            fundamental_matrix_model m;
            load(m, M18);
            if (thr > 0)
                m.inlier_threshold = thr;
            m.checkDegeneracy(c, inl);
            store(m, M18);
            for (size_t i = 0; i < n; i++)
                inliers[i] = inl[i];
        });
    }

    // essential_matrix_model::decompose -> 4 x (qx,qy,qz,qw, tx,ty,tz)
    void ocbh_decompose_essential(const double *M18, double *poses28)
    {
        essential_matrix_model m;
        load(m, M18);
        std::array<decomposed_pose, 4> poses;
        m.decompose({}, {}, poses);
        for (int i = 0; i < 4; i++)
        {
            for (int k = 0; k < 4; k++)
                poses28[7 * i + k] = poses[i].orientation.coeffs()[k];
            for (int k = 0; k < 3; k++)
                poses28[7 * i + 4 + k] = poses[i].position[k];
        }
    }

    // homography_model::decompose -> 4 x (qx,qy,qz,qw, tx,ty,tz, score); returns the reference's bool
    int ocbh_decompose_homography(const double *M18, const double *corr, size_t n, const uint8_t *inliers,
                                  double *poses32)
    {
        homography_model m;
        load(m, M18);
        const std::vector<correspondence> c = make_corr(corr, n);
        std::array<decomposed_pose, 4> poses;
        const bool ok = m.decompose(c, make_flags(inliers, n), poses);
        for (int i = 0; i < 4; i++)
        {
            for (int k = 0; k < 4; k++)
                poses32[8 * i + k] = poses[i].orientation.coeffs()[k];
            for (int k = 0; k < 3; k++)
                poses32[8 * i + 4 + k] = poses[i].position[k];
            poses32[8 * i + 7] = poses[i].score;
        }
        return ok ? 1 : 0;
    }
    int ocbh_decompose_homography_mat(const double *H9, double *R36, double *t12, double *n12)
    {
        return ocb_host::detail::decompose_homography_mat(H9, R36, t12, n12);
    }

    // assembleInliers -> rows of (pixel_1.x, pixel_1.y, pixel_2.x, pixel_2.y) + (idx1, idx2, match_index)
    size_t ocbh_assemble_inliers(const size_t *m_i1, const size_t *m_i2, const double *m_dist, size_t n_matches,
                                 const uint8_t *inliers, const double *xy1, size_t nf1, const double *xy2, size_t nf2,
                                 double *out_pixels4, size_t *out_idx3)
    {
        std::vector<feature_match> matches(n_matches);
        for (size_t i = 0; i < n_matches; i++)
            matches[i] = feature_match{m_i1[i], m_i2[i], m_dist[i]};
        const std::vector<feature_2d> f1 = make_features(xy1, nullptr, nullptr, nf1);
        const std::vector<feature_2d> f2 = make_features(xy2, nullptr, nullptr, nf2);
        std::vector<feature_match_denormalized> list;
        assembleInliers(matches, make_flags(inliers, n_matches), f1, f2, list);
        for (size_t i = 0; i < list.size(); i++)
        {
            out_pixels4[4 * i + 0] = list[i].pixel_1.x(), out_pixels4[4 * i + 1] = list[i].pixel_1.y();
            out_pixels4[4 * i + 2] = list[i].pixel_2.x(), out_pixels4[4 * i + 3] = list[i].pixel_2.y();
            out_idx3[3 * i + 0] = list[i].feature_index_1, out_idx3[3 * i + 1] = list[i].feature_index_2;
            out_idx3[3 * i + 2] = list[i].match_index;
        }
        return list.size();
    }

    // ---- linear algebra (host only) ------------------------------------------------------------------------
    void ocbh_full_piv_lu_solve(const double *A_colmajor, int rows, int cols, const double *b, double *x)
    {
        ocb_host::linalg::ColMat A(rows, cols);
        std::memcpy(A.a.data(), A_colmajor, sizeof(double) * rows * cols);
        const std::vector<double> r = ocb_host::linalg::full_piv_lu_solve(A, std::vector<double>(b, b + rows));
        std::memcpy(x, r.data(), sizeof(double) * cols);
    }
    void ocbh_invert3(const double *m, double *out) { ocb_host::linalg::invert3(m, out); }
    void ocbh_jacobi_svd(const double *A_colmajor, int n, double *U, double *S, double *V)
    {
        ocb_host::linalg::ColMat A(n, n);
        std::memcpy(A.a.data(), A_colmajor, sizeof(double) * n * n);
        const ocb_host::linalg::Svd s = ocb_host::linalg::jacobi_svd(A, true, true);
        std::memcpy(U, s.U.a.data(), sizeof(double) * n * n);
        std::memcpy(V, s.V.a.data(), sizeof(double) * n * n);
        std::memcpy(S, s.sigma.data(), sizeof(double) * n);
    }
    void ocbh_jacobi_svd_tall(const double *A_colmajor, int rows, int cols, double *S, double *V)
    {
        ocb_host::linalg::ColMat A(rows, cols);
        std::memcpy(A.a.data(), A_colmajor, sizeof(double) * rows * cols);
        const ocb_host::linalg::Svd s = ocb_host::linalg::jacobi_svd_tall(A);
        std::memcpy(V, s.V.a.data(), sizeof(double) * cols * cols);
        std::memcpy(S, s.sigma.data(), sizeof(double) * cols);
    }

    // ---- the reference's run_parallel shape (src/pipeline/pipeline.cpp:42-49): one closure per pair executed by
    // `threads` OpenMP workers, each closure = match_features_subset on its pair (link_stage.cpp:83-84). Every
    // worker drives its own CUDA stream through the thread-safe C ABI. Returns wall seconds.
    int ocbh_run_parallel_match(const uint64_t *q, const uint64_t *c, size_t n_pairs, size_t n1, size_t n2, int threads,
                                size_t *n_matches, double *seconds)
    {
        if (threads <= 0)
            threads = omp_get_num_procs();
        std::vector<std::vector<feature_2d>> fq(n_pairs), fc(n_pairs);
        for (size_t p = 0; p < n_pairs; p++)
        {
            fq[p] = make_features(nullptr, nullptr, q + p * n1 * 8, n1);
            fc[p] = make_features(nullptr, nullptr, c + p * n2 * 8, n2);
        }
        std::vector<size_t> idx1(n1), idx2(n2);
        for (size_t i = 0; i < n1; i++)
            idx1[i] = i;
        for (size_t i = 0; i < n2; i++)
            idx2[i] = i;
        size_t total = 0;
        int failed = 0;
        const auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads) reduction(+ : total)
        for (size_t p = 0; p < n_pairs; p++)
        {
            try
            {
                total += match_features_subset(fq[p], fc[p], idx1, idx2).size();
            }
            catch (const std::exception &e)
            {
#pragma omp critical
                {
                    t_err = e.what();
                    failed = 1;
                }
            }
        }
        const auto t1 = std::chrono::steady_clock::now();
        *n_matches = total;
        *seconds = std::chrono::duration<double>(t1 - t0).count();
        return failed ? -1 : 0;
    }
    // Same driver on feature sets that already live in host memory (ocbh_features_create), repeated `reps` times:
    // reps * n_pairs closures, each one match_features_subset (or its cross-checked variant) on all features of
    // its pair. Used by bench.py for the concurrent-callers end-to-end number.
    int ocbh_run_parallel_handles(const void *const *hq, const void *const *hc, size_t n_pairs, int threads,
                                  int cross_check, int reps, size_t *n_matches, double *seconds)
    {
        if (threads <= 0)
            threads = omp_get_num_procs();
        std::vector<std::vector<size_t>> idx1(n_pairs), idx2(n_pairs);
        for (size_t p = 0; p < n_pairs; p++)
        {
            idx1[p].resize(static_cast<const std::vector<feature_2d> *>(hq[p])->size());
            idx2[p].resize(static_cast<const std::vector<feature_2d> *>(hc[p])->size());
            for (size_t i = 0; i < idx1[p].size(); i++)
                idx1[p][i] = i;
            for (size_t i = 0; i < idx2[p].size(); i++)
                idx2[p][i] = i;
        }
        size_t total = 0;
        int failed = 0;
        const size_t jobs = n_pairs * (size_t)std::max(reps, 1);
        const auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads) reduction(+ : total)
        for (size_t job = 0; job < jobs; job++)
        {
            const size_t p = job % n_pairs;
            const auto &f1 = *static_cast<const std::vector<feature_2d> *>(hq[p]);
            const auto &f2 = *static_cast<const std::vector<feature_2d> *>(hc[p]);
            try
            {
                if (cross_check)
                {
                    std::vector<bool> mutual;
                    total += ocb_host::match_features_subset_cross_checked(f1, f2, idx1[p], idx2[p], mutual).size();
                }
                else
                    total += match_features_subset(f1, f2, idx1[p], idx2[p]).size();
            }
            catch (const std::exception &e)
            {
#pragma omp critical
                {
                    t_err = e.what();
                    failed = 1;
                }
            }
        }
        const auto t1 = std::chrono::steady_clock::now();
        *n_matches = total;
        *seconds = std::chrono::duration<double>(t1 - t0).count();
        return failed ? -1 : 0;
    }
    // ---- distort_keypoints / image_to_3d: cam8 = (f, ppx, ppy, k1, k2, k3, p1, p2); rays out [n][3]
    void ocbh_image_to_3d(const double *xy, size_t n, const double *cam8, double *rays)
    {
        DifferentiableCameraModel<double> m;
        m.focal_length_pixels = cam8[0];
        m.principle_point = Eigen::Vector2d(cam8[1], cam8[2]);
        m.radial_distortion = Eigen::Vector3d(cam8[3], cam8[4], cam8[5]);
        m.tangential_distortion = Eigen::Vector2d(cam8[6], cam8[7]);
        for (size_t i = 0; i < n; i++)
        {
            const Eigen::Vector3d r = image_to_3d(Eigen::Vector2d(xy[2 * i], xy[2 * i + 1]), m);
            rays[3 * i] = r[0], rays[3 * i + 1] = r[1], rays[3 * i + 2] = r[2];
        }
    }

    // ---- batched LinkStage runner (link_batch.hpp). images: handles from ocbh_features_create, num_sparse[i],
    // cam8[i]; pairs: [n_pairs][2] image indices. Returns an opaque result handle (nullptr on error).
    struct LinkResultHandle
    {
        std::vector<camera_relations> relations;
        ocb_host::LinkStats stats;
    };
    // device_tail: 1 = ratio test, compaction and rays on the device (default), 0 = on the host. n_devices > 1: the pair
    // list is partitioned over that many GPUs of this process (positions2: [n_images][2], nullable). packed_*: see
    // LinkOptions (nullable; single-device runs only).
    void *ocbh_link_pairs(const void *const *image_handles, const size_t *num_sparse, const double *cam8, size_t n_images,
                          const size_t *pairs2, size_t n_pairs, int threads, size_t pairs_per_submission, int run_ransac,
                          double spacing, int device_tail, int n_devices, const double *positions2, uint32_t *packed_out,
                          size_t packed_capacity, uint64_t *packed_offsets, uint64_t *packed_counts)
    {
        auto *res = new LinkResultHandle;
        const int rc = guarded([&] {
            std::vector<ocb_host::LinkImage> images(n_images);
            for (size_t i = 0; i < n_images; i++)
            {
                images[i].features = static_cast<const std::vector<feature_2d> *>(image_handles[i]);
                images[i].num_sparse_features = num_sparse ? num_sparse[i] : 0;
                const double *c = cam8 + 8 * i;
                images[i].model.focal_length_pixels = c[0];
                images[i].model.principle_point = Eigen::Vector2d(c[1], c[2]);
                images[i].model.radial_distortion = Eigen::Vector3d(c[3], c[4], c[5]);
                images[i].model.tangential_distortion = Eigen::Vector2d(c[6], c[7]);
            }
            std::vector<ocb_host::LinkPair> pairs(n_pairs);
            for (size_t p = 0; p < n_pairs; p++)
                pairs[p] = ocb_host::LinkPair{pairs2[2 * p], pairs2[2 * p + 1]};
            ocb_host::LinkOptions opt;
            opt.threads = threads;
            if (pairs_per_submission)
                opt.pairs_per_submission = pairs_per_submission;
            opt.run_ransac = run_ransac != 0;
            if (spacing > 0)
                opt.coarse_spacing_pixels = spacing;
            opt.device_tail = device_tail != 0;
            opt.packed_out = packed_out, opt.packed_capacity = packed_capacity;
            opt.packed_offsets = packed_offsets, opt.packed_counts = packed_counts;
            if (n_devices > 1)
                res->relations = ocb_host::link_pairs_multi(images, pairs, positions2, n_devices, opt, &res->stats);
            else
                res->relations = ocb_host::link_pairs(images, pairs, opt, &res->stats);
        });
        if (rc)
        {
            delete res;
            return nullptr;
        }
        return res;
    }
    void ocbh_link_free(void *h) { delete static_cast<LinkResultHandle *>(h); }
    // stats9: seconds subsample+upload, gpu match, tail, total, comparisons, matches, ransac inliers, seconds setup,
    // seconds release
    void ocbh_link_stats(const void *h, double *stats8)
    {
        const auto &s = static_cast<const LinkResultHandle *>(h)->stats;
        stats8[0] = s.seconds_subsample_upload, stats8[1] = s.seconds_match_gpu, stats8[2] = s.seconds_tail;
        stats8[3] = s.seconds_total, stats8[4] = (double)s.comparisons, stats8[5] = (double)s.matches;
        stats8[6] = (double)s.ransac_inliers, stats8[7] = s.seconds_setup + 1e3 * 0;
        stats8[8] = s.seconds_release;
    }
    void ocbh_link_sizes(const void *h, size_t p, size_t *n_matches, size_t *n_inlier_matches)
    {
        const auto &r = static_cast<const LinkResultHandle *>(h)->relations[p];
        *n_matches = r.matches.size();
        *n_inlier_matches = r.inlier_matches.size();
    }
    // one pair: matches (i1, i2, dist), H9 column-major, relation type, poses [4][8], inlier pixels4 + idx3
    void ocbh_link_get(const void *h, size_t p, size_t *m_i1, size_t *m_i2, double *m_dist, double *H9, int *relation_type,
                       double *poses32, double *inl_pixels4, size_t *inl_idx3)
    {
        const auto &r = static_cast<const LinkResultHandle *>(h)->relations[p];
        for (size_t i = 0; i < r.matches.size(); i++)
            m_i1[i] = r.matches[i].feature_index_1, m_i2[i] = r.matches[i].feature_index_2,
            m_dist[i] = r.matches[i].distance;
        std::memcpy(H9, r.ransac_relation.data(), 72);
        *relation_type = (int)r.relationType;
        for (int i = 0; i < 4; i++)
        {
            for (int k = 0; k < 4; k++)
                poses32[8 * i + k] = r.relative_poses[i].orientation.coeffs()[k];
            for (int k = 0; k < 3; k++)
                poses32[8 * i + 4 + k] = r.relative_poses[i].position[k];
            poses32[8 * i + 7] = r.relative_poses[i].score;
        }
        for (size_t i = 0; i < r.inlier_matches.size(); i++)
        {
            const auto &m = r.inlier_matches[i];
            inl_pixels4[4 * i] = m.pixel_1[0], inl_pixels4[4 * i + 1] = m.pixel_1[1];
            inl_pixels4[4 * i + 2] = m.pixel_2[0], inl_pixels4[4 * i + 3] = m.pixel_2[1];
            inl_idx3[3 * i] = m.feature_index_1, inl_idx3[3 * i + 1] = m.feature_index_2, inl_idx3[3 * i + 2] = m.match_index;
        }
    }
    // All match lists of a link result as one flat array of 12-byte records {feature_index_1, feature_index_2, integer
    // Hamming distance} (distance = d * (1.0 / 486) exactly), pair after pair in pair order; counts[p] = matches of pair
    // p. records == nullptr: only the counts. This is what a rank contributes to the host gather of the match lists.
    size_t ocbh_link_pack_matches(const void *h, uint64_t *counts, uint32_t *records, int threads)
    {
        const auto &rel = static_cast<const LinkResultHandle *>(h)->relations;
        std::vector<size_t> off(rel.size() + 1, 0);
        for (size_t p = 0; p < rel.size(); p++)
        {
            counts[p] = rel[p].matches.size();
            off[p + 1] = off[p] + rel[p].matches.size();
        }
        if (records)
        {
#pragma omp parallel for schedule(dynamic, 16) num_threads(threads > 0 ? threads : omp_get_num_procs())
            for (size_t p = 0; p < rel.size(); p++)
            {
                uint32_t *out = records + 3 * off[p];
                for (const feature_match &m : rel[p].matches)
                {
                    out[0] = (uint32_t)m.feature_index_1, out[1] = (uint32_t)m.feature_index_2;
                    out[2] = (uint32_t)std::lround(m.distance * feature_2d::DESCRIPTOR_BITS);
                    out += 3;
                }
            }
        }
        return off[rel.size()];
    }

    // ---- overlap-graph partition (partition.hpp): owner[i] = part of image i, pair_part[k] = part of pair k,
    // halo[part * n_images + i] = 1 when image i is a halo image of that part
    int ocbh_partition_pairs(const double *positions2, size_t n_images, const size_t *pairs2, size_t n_pairs, size_t world,
                             uint32_t *owner, uint32_t *pair_part, uint8_t *halo)
    {
        return guarded([&] {
            std::vector<ocb_host::LinkPair> pairs(n_pairs);
            for (size_t p = 0; p < n_pairs; p++)
                pairs[p] = ocb_host::LinkPair{pairs2[2 * p], pairs2[2 * p + 1]};
            const std::vector<ocb_host::PairShard> shards = ocb_host::partition_pairs(positions2, n_images, pairs, world);
            std::memset(halo, 0, world * n_images);
            for (size_t r = 0; r < world; r++)
            {
                for (size_t i : shards[r].owned_images)
                    owner[i] = (uint32_t)r;
                for (size_t i : shards[r].halo_images)
                    halo[r * n_images + i] = 1;
                for (size_t k : shards[r].pair_ids)
                    pair_part[k] = (uint32_t)r;
            }
        });
    }
    uint32_t ocbh_hilbert_index(int order, int x, int y) { return ocb_host::hilbert_index(order, x, y); }
    void ocbh_hilbert_order(const double *positions2, size_t n, size_t *order_out)
    {
        const std::vector<size_t> o = ocb_host::hilbert_order(positions2, n);
        std::memcpy(order_out, o.data(), n * sizeof(size_t));
    }

    // the host's own PROSAC ordering (ransac.cpp:83-90: std::sort of 0 .. n-1 by quality, ascending) for a quality array
    void ocbh_prosac_order(const double *quality, size_t n, uint32_t *order)
    {
        struct Ranked
        {
            double quality;
            size_t idx;
        };
        std::vector<Ranked> ranked(n);
        for (size_t i = 0; i < n; i++)
            ranked[i] = Ranked{quality[i], i};
        std::sort(ranked.begin(), ranked.end(), [](const Ranked &a, const Ranked &b) { return a.quality < b.quality; });
        for (size_t i = 0; i < n; i++)
            order[i] = (uint32_t)ranked[i].idx;
    }

    // refit_evaluate_batch over n_jobs correspondence sets (rows offsets[j] .. offsets[j + 1] of corr); inl: in = the
    // previous inliers, out = the last evaluate's; M18 [n_jobs][18] out; thr: inlier threshold (<= 0: the model's default)
    int ocbh_refit_evaluate_batch(const double *corr, const size_t *offsets, size_t n_jobs, int rounds, double thr,
                                  double *scores, double *M18, uint8_t *inl)
    {
        return guarded([&] {
            std::vector<std::vector<correspondence>> corrs(n_jobs);
            std::vector<homography_model> models(n_jobs);
            std::vector<std::vector<bool>> inliers(n_jobs);
            std::vector<ocb_host::RefitJob> jobs(n_jobs);
            for (size_t j = 0; j < n_jobs; j++)
            {
                const size_t n = offsets[j + 1] - offsets[j];
                corrs[j] = make_corr(corr + offsets[j] * 7, n);
                inliers[j].resize(n);
                for (size_t k = 0; k < n; k++)
                    inliers[j][k] = inl[offsets[j] + k] != 0;
                if (thr > 0)
                    models[j].inlier_threshold = thr;
                jobs[j].matches = &corrs[j], jobs[j].model = &models[j], jobs[j].inliers = &inliers[j];
            }
            ocb_host::refit_evaluate_batch(jobs, rounds);
            for (size_t j = 0; j < n_jobs; j++)
            {
                scores[j] = jobs[j].score;
                ocb_host::detail::pack_model(models[j], M18 + j * 18);
                for (size_t k = 0; k < inliers[j].size(); k++)
                    inl[offsets[j] + k] = inliers[j][k] ? 1 : 0;
            }
        });
    }

    int ocbh_ransac_batch(int kind, const double *corr, const size_t *offsets, size_t n_jobs, int threads, double *scores,
                          double *M18, uint8_t *inl, size_t *stats2)
    {
        return guarded([&] {
            if (kind == OCB_MODEL_HOMOGRAPHY)
                run_batch<homography_model>(corr, offsets, n_jobs, threads, scores, M18, inl, stats2);
            else if (kind == OCB_MODEL_ESSENTIAL)
                run_batch<essential_matrix_model>(corr, offsets, n_jobs, threads, scores, M18, inl, stats2);
            else
                run_batch<fundamental_matrix_model>(corr, offsets, n_jobs, threads, scores, M18, inl, stats2);
        });
    }
}
