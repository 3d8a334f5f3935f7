// TEST INFRASTRUCTURE ONLY -- flat C entry points over oc_oracle.hpp for ctypes (tests/, smoke(),
// bench.py's cpu_baseline / --impl reference legs). Not linked by the product.
#include "oc_oracle.hpp"

#include <algorithm>
#include <chrono>
#include <cstring>
#include <omp.h>

using namespace oc_oracle;

namespace
{
Model make_model(int kind, const double *M18, double thr)
{
    Model m(kind);
    if (thr > 0)
        m.thr = thr;
    if (M18)
    {
        std::memcpy(m.M, M18, sizeof(double) * 9);
        std::memcpy(m.Minv, M18 + 9, sizeof(double) * 9);
    }
    return m;
}
void store_model(const Model &m, double *M18)
{
    std::memcpy(M18, m.M, sizeof(double) * 9);
    std::memcpy(M18 + 9, m.Minv, sizeof(double) * 9);
}
} // namespace

extern "C"
{
    void oco_match_top2(const uint64_t *q, size_t n1, const uint64_t *c, size_t n2, uint32_t *best_k,
                        uint16_t *best_d, uint16_t *second_d)
    {
        match_top2(q, n1, c, n2, best_k, best_d, second_d);
    }
    void oco_match_col_best(const uint64_t *q, size_t n1, const uint64_t *c, size_t n2, uint32_t *col_best_q)
    {
        match_col_best(q, n1, c, n2, col_best_q);
    }
    // -> per list: best position, best / second distance (double), accepted flag (dense_stereo.cpp:251-276)
    void oco_match_lists(const uint64_t *q, const uint64_t *c, const uint32_t *list_query, const uint64_t *list_begin,
                         const uint32_t *list_candidates, size_t n_lists, uint32_t *best_pos, double *best_dist,
                         double *second_dist, uint8_t *good)
    {
        match_lists_top2(q, c, list_query, list_begin, list_candidates, n_lists, best_pos, best_dist, second_dist);
        for (size_t l = 0; l < n_lists; l++)
            good[l] = guided_good_match(list_begin[l + 1] - list_begin[l], best_dist[l], second_dist[l]);
    }
    // the same loop timed under OpenMP over lists (the reference parallelises over source images, :174-175)
    double oco_bench_match_lists(const uint64_t *q, const uint64_t *c, const uint32_t *list_query,
                                 const uint64_t *list_begin, const uint32_t *list_candidates, size_t n_lists,
                                 int threads, uint32_t *best_pos, double *best_dist, double *second_dist)
    {
        if (threads <= 0)
            threads = omp_get_num_procs();
        const size_t chunk = 256;
        const long chunks = (long)((n_lists + chunk - 1) / chunk);
        auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel for schedule(dynamic) num_threads(threads)
        for (long b = 0; b < chunks; b++)
        {
            const size_t l0 = (size_t)b * chunk, l1 = std::min(n_lists, l0 + chunk);
            // list_begin offsets are absolute, so a sub-range is the same call on shifted pointers
            match_lists_top2(q, c, list_query + l0, list_begin + l0, list_candidates, l1 - l0, best_pos + l0,
                             best_dist + l0, second_dist + l0);
        }
        return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }
    size_t oco_match_features_subset(const uint64_t *desc1, const uint64_t *desc2, const size_t *idx1, size_t n1,
                                     const size_t *idx2, size_t n2, size_t *out_i1, size_t *out_i2, double *out_dist)
    {
        std::vector<Match> r = match_features_subset(desc1, desc2, idx1, n1, idx2, n2);
        for (size_t i = 0; i < r.size(); i++)
        {
            out_i1[i] = r[i].feature_index_1;
            out_i2[i] = r[i].feature_index_2;
            out_dist[i] = r[i].distance;
        }
        return r.size();
    }
    size_t oco_subsample(const double *xy, const float *strength, size_t n, double spacing, size_t count,
                         size_t *out_idx)
    {
        std::vector<size_t> r = spatially_subsample_feature_indices(xy, strength, n, spacing, count);
        std::memcpy(out_idx, r.data(), r.size() * sizeof(size_t));
        return r.size();
    }
    double oco_error(int kind, const double *M18, double thr, const double *corr7)
    {
        return error(make_model(kind, M18, thr), *reinterpret_cast<const Corr *>(corr7));
    }
    double oco_evaluate(int kind, const double *M18, double thr, const double *corr, size_t n, uint8_t *inliers)
    {
        std::vector<bool> inl;
        double s = evaluate(make_model(kind, M18, thr), reinterpret_cast<const Corr *>(corr), n, inl);
        for (size_t i = 0; i < n; i++)
            inliers[i] = inl[i];
        return s;
    }
    void oco_fit(int kind, const double *corr, const size_t *sample, double *M18)
    {
        Model m(kind);
        fit(m, reinterpret_cast<const Corr *>(corr), sample);
        store_model(m, M18);
    }
    void oco_fit_inliers(int kind, double *M18, const double *corr, size_t n, const uint8_t *inliers)
    {
        Model m = make_model(kind, M18, 0);
        std::vector<bool> inl(n);
        for (size_t i = 0; i < n; i++)
            inl[i] = inliers[i] != 0;
        fit_inliers(m, reinterpret_cast<const Corr *>(corr), n, inl);
        store_model(m, M18);
    }
    int oco_check_sample_degeneracy_h(const double *corr, const size_t *sample)
    {
        return check_sample_degeneracy_h(reinterpret_cast<const Corr *>(corr), sample) ? 1 : 0;
    }
    void oco_check_degeneracy_f(double *M18, double thr, const double *corr, size_t n, uint8_t *inliers)
    {
        Model m = make_model(MODEL_FUNDAMENTAL, M18, thr);
        std::vector<bool> inl(n);
        for (size_t i = 0; i < n; i++)
            inl[i] = inliers[i] != 0;
        check_degeneracy_f(m, reinterpret_cast<const Corr *>(corr), n, inl);
        for (size_t i = 0; i < n; i++)
            inliers[i] = inl[i];
        store_model(m, M18);
    }
    double oco_ransac(int kind, const double *corr, size_t n, double *M18, uint8_t *inliers, size_t *trace4)
    {
        Model m(kind);
        std::vector<bool> inl;
        RansacTrace tr;
        double s = ransac(reinterpret_cast<const Corr *>(corr), n, m, inl, &tr);
        for (size_t i = 0; i < n; i++)
            inliers[i] = inl[i];
        store_model(m, M18);
        if (trace4)
        {
            trace4[0] = tr.iterations;
            trace4[1] = tr.improvements;
            trace4[2] = tr.rejected;
            trace4[3] = tr.degenerate;
        }
        return s;
    }
    // returns 0 when n < MINIMUM_POINTS (no stream), else 1
    int oco_hypothesis_stream(int kind, const double *corr, size_t n, size_t count, size_t *eval_order,
                              size_t *samples)
    {
        std::vector<size_t> eo, sm;
        hypothesis_stream(reinterpret_cast<const Corr *>(corr), n, kind, count, eo, sm);
        if (eo.empty())
            return 0;
        std::memcpy(eval_order, eo.data(), eo.size() * sizeof(size_t));
        std::memcpy(samples, sm.data(), sm.size() * sizeof(size_t));
        return 1;
    }
    void oco_score_hypotheses(int kind, const double *M18s, size_t h, double thr, const double *corr, size_t n,
                              const size_t *order, double *score, uint32_t *count, uint32_t *bits, int threads)
    {
        const size_t words = (n + 31) / 32;
        if (threads <= 0)
            threads = omp_get_num_procs();
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
        for (size_t i = 0; i < h; i++)
        {
            Model m = make_model(kind, M18s + 18 * i, thr);
            score_hypothesis(m, reinterpret_cast<const Corr *>(corr), n, order, &score[i], &count[i],
                             bits ? bits + i * words : nullptr);
        }
    }
    void oco_fullpivlu_solve(const double *A, int rows, int cols, const double *b, double *x)
    {
        la::fullpivlu_solve(A, rows, cols, b, x);
    }
    void oco_inverse3(const double *M, double *Minv) { la::inverse3(M, Minv); }
    void oco_jacobi_svd_square(const double *A, int n, double *U, double *S, double *V)
    {
        la::jacobi_svd_square(A, n, U, S, V);
    }
    void oco_jacobi_svd_tall_v(const double *A, int rows, int cols, double *S, double *V)
    {
        la::jacobi_svd_tall_v(A, rows, cols, S, V);
    }

    // CPU-baseline timing leg: match `n_pairs` independent pairs the way the reference's run_parallel
    // does (src/pipeline/pipeline.cpp:42-49: omp parallel for schedule(dynamic,1) over closures, one pair
    // per thread). Pair p matches rows q[p] (n1 each) against rows c[p] (n2 each), identity indices.
    // Returns wall seconds; *n_matches receives the total number of ratio-test survivors.
    double oco_bench_match_pairs(const uint64_t *q, const uint64_t *c, size_t n_pairs, size_t n1, size_t n2,
                                 int threads, size_t *n_matches)
    {
        if (threads <= 0)
            threads = omp_get_num_procs();
        std::vector<size_t> idx1(n1), idx2(n2);
        for (size_t i = 0; i < n1; i++)
            idx1[i] = i;
        for (size_t i = 0; i < n2; i++)
            idx2[i] = i;
        size_t total = 0;
        auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads) reduction(+ : total)
        for (size_t p = 0; p < n_pairs; p++)
        {
            std::vector<Match> r = match_features_subset(q + p * n1 * 8, c + p * n2 * 8, idx1.data(), n1,
                                                         idx2.data(), n2);
            total += r.size();
        }
        auto t1 = std::chrono::steady_clock::now();
        if (n_matches)
            *n_matches = total;
        return std::chrono::duration<double>(t1 - t0).count();
    }
    int oco_num_procs() { return omp_get_num_procs(); }
}

// ---- synthetic scenes of test/test_ransac_benchmark.cpp:12-116 (same std::mt19937 /
// uniform_real_distribution calls, so the random draws are the reference test's draws) -------------
#include <cmath>
#include <random>
namespace
{
void matvec3(const double *M /*row-major*/, const double *v, double *out)
{
    for (int r = 0; r < 3; r++)
        out[r] = (M[3 * r] * v[0] + M[3 * r + 1] * v[1]) + M[3 * r + 2] * v[2];
}
void gt_homography(double *G /*row-major*/)
{
    // R = AngleAxis(0.1, Z); t = (0.05,-0.03,0); n = (0,0,1); G = R + t n^T / 10; G /= G(2,2)
    const double c = std::cos(0.1), s = std::sin(0.1);
    const double R[9] = {c, -s, 0, s, c, 0, 0, 0, 1};
    const double t[3] = {0.05, -0.03, 0.0}, n[3] = {0, 0, 1};
    for (int r = 0; r < 3; r++)
        for (int k = 0; k < 3; k++)
            G[3 * r + k] = R[3 * r + k] + t[r] * n[k] / 10.0;
    const double g22 = G[8];
    for (int i = 0; i < 9; i++)
        G[i] /= g22;
}
void put(double *corr, size_t i, const double *a, const double *b)
{
    for (int k = 0; k < 3; k++)
    {
        corr[7 * i + k] = a[k];
        corr[7 * i + 3 + k] = b[k];
    }
    corr[7 * i + 6] = 0;
}
} // namespace

extern "C"
{
    // SyntheticScene::homography (test_ransac_benchmark.cpp:18-58). gt9 row-major.
    void oco_scene_homography(size_t n_inliers, size_t n_outliers, unsigned seed, double *corr, double *gt9)
    {
        std::mt19937 rng(seed);
        std::uniform_real_distribution<double> point_dist(-1.0, 1.0);
        std::uniform_real_distribution<double> outlier_dist(-2.0, 2.0);
        gt_homography(gt9);
        for (size_t i = 0; i < n_inliers; i++)
        {
            double p1[3];
            p1[0] = point_dist(rng);
            p1[1] = point_dist(rng);
            p1[2] = 1.0;
            double p2[3];
            matvec3(gt9, p1, p2);
            const double z = p2[2];
            for (double &v : p2)
                v /= z;
            put(corr, i, p1, p2);
        }
        for (size_t i = 0; i < n_outliers; i++)
        {
            double a[3], b[3];
            a[0] = outlier_dist(rng);
            a[1] = outlier_dist(rng);
            a[2] = 1.0;
            b[0] = outlier_dist(rng);
            b[1] = outlier_dist(rng);
            b[2] = 1.0;
            put(corr, n_inliers + i, a, b);
        }
    }
    // homography_near_degenerate scene (test_ransac_benchmark.cpp:226-265): 20 near-collinear + 80 spread.
    void oco_scene_homography_near_degenerate(double *corr, double *gt9)
    {
        std::mt19937 rng(42);
        std::uniform_real_distribution<double> noise(-0.001, 0.001);
        gt_homography(gt9);
        size_t at = 0;
        for (int i = 0; i < 20; i++)
        {
            const double t_param = -1.0 + 2.0 * i / 19.0;
            double p1[3] = {t_param, 0.5 + noise(rng), 1.0}, p2[3];
            matvec3(gt9, p1, p2);
            const double z = p2[2];
            for (double &v : p2)
                v /= z;
            put(corr, at++, p1, p2);
        }
        std::uniform_real_distribution<double> point_dist(-1.0, 1.0);
        for (int i = 0; i < 80; i++)
        {
            double p1[3];
            p1[0] = point_dist(rng);
            p1[1] = point_dist(rng);
            p1[2] = 1.0;
            double p2[3];
            matvec3(gt9, p1, p2);
            const double z = p2[2];
            for (double &v : p2)
                v /= z;
            put(corr, at++, p1, p2);
        }
    }
    // SyntheticScene::fundamental (test_ransac_benchmark.cpp:60-116). gt9 row-major, Frobenius-normalised.
    void oco_scene_fundamental(size_t n_inliers, size_t n_outliers, double planar_fraction, unsigned seed,
                               double *corr, double *gt9)
    {
        std::mt19937 rng(seed);
        std::uniform_real_distribution<double> xy_dist(-1.0, 1.0);
        std::uniform_real_distribution<double> z_dist(5.0, 15.0);
        std::uniform_real_distribution<double> outlier_dist(-1.0, 1.0);
        const double c = std::cos(0.15), s = std::sin(0.15);
        const double R2[9] = {c, 0, s, 0, 1, 0, -s, 0, c};
        const double t2[3] = {0.5, 0.0, 0.0};
        const double mt2[3] = {-t2[0], -t2[1], -t2[2]};
        double e2[3];
        matvec3(R2, mt2, e2);
        const double ex[9] = {0, -e2[2], e2[1], e2[2], 0, -e2[0], -e2[1], e2[0], 0};
        double nrm = 0;
        for (int r = 0; r < 3; r++)
            for (int k = 0; k < 3; k++)
            {
                double v = 0;
                for (int j = 0; j < 3; j++)
                    v += ex[3 * r + j] * R2[3 * j + k];
                gt9[3 * r + k] = v;
                nrm += v * v;
            }
        nrm = std::sqrt(nrm);
        for (int i = 0; i < 9; i++)
            gt9[i] /= nrm;
        const size_t n_planar = static_cast<size_t>(n_inliers * planar_fraction);
        for (size_t i = 0; i < n_inliers; i++)
        {
            double X[3];
            if (i < n_planar)
            {
                X[0] = xy_dist(rng) * 3;
                X[1] = xy_dist(rng) * 3;
                X[2] = 10.0;
            }
            else
            {
                X[0] = xy_dist(rng) * 3;
                X[1] = xy_dist(rng) * 3;
                X[2] = z_dist(rng);
            }
            double x1[3] = {X[0], X[1], X[2]};
            const double d[3] = {X[0] - t2[0], X[1] - t2[1], X[2] - t2[2]};
            double x2[3];
            matvec3(R2, d, x2);
            const double z1 = x1[2], z2 = x2[2];
            for (double &v : x1)
                v /= z1;
            for (double &v : x2)
                v /= z2;
            put(corr, i, x1, x2);
        }
        for (size_t i = 0; i < n_outliers; i++)
        {
            double a[3], b[3];
            a[0] = outlier_dist(rng);
            a[1] = outlier_dist(rng);
            a[2] = 1.0;
            b[0] = outlier_dist(rng);
            b[1] = outlier_dist(rng);
            b[2] = 1.0;
            put(corr, n_inliers + i, a, b);
        }
    }
}
