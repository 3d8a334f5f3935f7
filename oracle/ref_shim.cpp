// TEST INFRASTRUCTURE ONLY -- flat C entry points over the REFERENCE'S OWN object code
// (src/match/match_features.cpp and src/model_inliers/ransac.cpp, compiled in place from /root/reference
// by oracle/Makefile into oracle/_ref/). Used to pin the oracle restatement (tests/golden generation,
// tests on this container) and as the "reference" CPU arm of bench.py. Never linked by the product.
#include <opencalibration/match/match_features.hpp>
#include <opencalibration/model_inliers/ransac.hpp>

#include <jk/KDTree.h>

#include <chrono>
#include <limits>
#include <cstddef>
#include <cstring>
#include <omp.h>

using namespace opencalibration;

static_assert(sizeof(std::bitset<feature_2d::DESCRIPTOR_BITS>) == 64, "bitset<486> is 8 x u64");

namespace
{
std::vector<feature_2d> make_features(const double *xy, const float *strength, const uint64_t *desc, size_t n)
{
    std::vector<feature_2d> f(n);
    for (size_t i = 0; i < n; i++)
    {
        if (xy)
        {
            f[i].location.x() = xy[2 * i];
            f[i].location.y() = xy[2 * i + 1];
        }
        if (strength)
            f[i].strength = strength[i];
        if (desc) // memory image of std::bitset<486> = the 8 little-endian u64 words
            std::memcpy(static_cast<void *>(&f[i].descriptor), desc + 8 * i, 64);
    }
    return f;
}
} // namespace

extern "C"
{
    size_t ocr_sizeof_feature_2d() { return sizeof(feature_2d); }
    size_t ocr_offsetof_descriptor() { return offsetof(feature_2d, descriptor); }
    size_t ocr_sizeof_feature_match() { return sizeof(feature_match); }
    size_t ocr_sizeof_correspondence() { return sizeof(correspondence); }

    size_t ocr_match_features_subset(const uint64_t *desc1, size_t nf1, const uint64_t *desc2, size_t nf2,
                                     const size_t *idx1, size_t n1, const size_t *idx2, size_t n2, size_t *out_i1,
                                     size_t *out_i2, double *out_dist)
    {
        std::vector<feature_2d> f1 = make_features(nullptr, nullptr, desc1, nf1);
        std::vector<feature_2d> f2 = make_features(nullptr, nullptr, desc2, nf2);
        std::vector<size_t> i1(idx1, idx1 + n1), i2(idx2, idx2 + n2);
        std::vector<feature_match> r = match_features_subset(f1, f2, i1, i2);
        for (size_t i = 0; i < r.size(); i++)
        {
            out_i1[i] = r[i].feature_index_1;
            out_i2[i] = r[i].feature_index_2;
            out_dist[i] = r[i].distance;
        }
        return r.size();
    }
    // The candidate lists of the dense stage exactly as the reference builds them (src/dense/dense_stereo.cpp:
    // 127-131 tree of an image's features, :244-246 radius search around each predicted position) with the
    // reference's own vendored jk-tree: list l = payloads of searcher.search(pred[l], radius_sq, max) in result
    // order. Two passes: out_nearby == NULL counts (returns the total), otherwise fills begin[n_q + 1] and nearby.
    size_t ocr_radius_lists(const double *cand_xy, size_t n_c, const double *pred_xy, size_t n_q, double radius_sq,
                            uint64_t *out_begin, uint32_t *out_nearby)
    {
        jk::tree::KDTree<size_t, 2, 8> tree;
        for (size_t i = 0; i < n_c; i++)
            tree.addPoint({cand_xy[2 * i], cand_xy[2 * i + 1]}, i);
        auto searcher = tree.searcher();
        size_t total = 0;
        for (size_t l = 0; l < n_q; l++)
        {
            const auto &nearby =
                searcher.search({pred_xy[2 * l], pred_xy[2 * l + 1]}, radius_sq, std::numeric_limits<size_t>::max());
            if (out_nearby)
            {
                out_begin[l] = total;
                for (size_t k = 0; k < nearby.size(); k++)
                    out_nearby[total + k] = (uint32_t)nearby[k].payload;
            }
            total += nearby.size();
        }
        if (out_nearby)
            out_begin[n_q] = total;
        return total;
    }
    size_t ocr_subsample(const double *xy, const float *strength, size_t n, double spacing, size_t count,
                         size_t *out_idx)
    {
        std::vector<feature_2d> f = make_features(xy, strength, nullptr, n);
        std::vector<size_t> r = spatially_subsample_feature_indices(f, spacing, count);
        std::memcpy(out_idx, r.data(), r.size() * sizeof(size_t));
        return r.size();
    }
    // kind: 0 homography, 1 essential, 2 fundamental. M18 = matrix (9, column-major) + inverse (H only).
    double ocr_ransac(int kind, const double *corr, size_t n, double *M18, uint8_t *inliers)
    {
        std::vector<correspondence> c(n);
        if (n)
            std::memcpy(static_cast<void *>(c.data()), corr, n * sizeof(correspondence));
        std::vector<bool> inl;
        double s = 0;
        for (int i = 0; i < 18; i++)
            M18[i] = NAN;
        if (kind == 0)
        {
            homography_model m;
            s = ransac(c, m, inl);
            std::memcpy(M18, m.homography.d, 72);
            std::memcpy(M18 + 9, m.homography_inverse.d, 72);
        }
        else if (kind == 1)
        {
            essential_matrix_model m;
            s = ransac(c, m, inl);
            std::memcpy(M18, m.essential_matrix.d, 72);
        }
        else
        {
            fundamental_matrix_model m;
            s = ransac(c, m, inl);
            std::memcpy(M18, m.fundamental_matrix.d, 72);
        }
        for (size_t i = 0; i < inl.size(); i++)
            inliers[i] = inl[i];
        return s;
    }

    // The reference's CPU path as the product runs it: one pair per OpenMP thread, schedule(dynamic,1)
    // (src/pipeline/pipeline.cpp:42-49). Returns wall seconds for n_pairs pairs of n1 x n2.
    double ocr_bench_match_pairs(const uint64_t *q, const uint64_t *c, size_t n_pairs, size_t n1, size_t n2,
                                 int threads, size_t *n_matches)
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
        auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads) reduction(+ : total)
        for (size_t p = 0; p < n_pairs; p++)
        {
            std::vector<feature_match> r = match_features_subset(fq[p], fc[p], idx1, idx2);
            total += r.size();
        }
        auto t1 = std::chrono::steady_clock::now();
        if (n_matches)
            *n_matches = total;
        return std::chrono::duration<double>(t1 - t0).count();
    }
    int ocr_num_procs() { return omp_get_num_procs(); }
}
