// TEST INFRASTRUCTURE ONLY.  CPU oracle: a restatement, in plain C++17 + libstdc++, of the
// reference's matching + RANSAC-scoring hot path (jkflying/opencalibration src/match and
// src/model_inliers).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load this; the product (opencalibration_b200/) never links it.
//
// Every function cites the reference file:line it restates (paths relative to the reference
// root).  Pinning status (see DESIGN.md "Oracle"):
//   * match / subsample / RANSAC driver (RNG, sampling, SPRT, LO, termination): PINNED --
//     checked bit-for-bit against the reference's own object code (oracle/_ref, built from
//     src/match/match_features.cpp and src/model_inliers/ransac.cpp in place) and against the
//     reference's KATs (test/test_match.cpp:90-107).
//   * residuals (error/evaluate) and fits (fit/fitInliers/checkDegeneracy): the arithmetic lives in
//     Eigen 3.4.0, which is neither vendored by the reference nor installed here, so it is restated
//     from Eigen's published algorithms (FullPivLU, 3x3 cofactor inverse, two-sided JacobiSVD) in a
//     documented canonical operation order: "parity unpinned" at the bit level, pinned to the
//     tolerances of test/test_ransac_unit.cpp and test/test_ransac_benchmark.cpp.
#pragma once
#include <array>
#include <cstddef>
#include <cstdint>
#include <vector>

namespace oc_oracle
{

// ---- layouts (mirror the reference structs) -------------------------------------------------
// include/opencalibration/types/correspondence.hpp:8-13 -- 7 doubles, 56 B.
struct Corr
{
    double m1[3];
    double m2[3];
    double quality;
};
// include/opencalibration/types/feature_match.hpp:10-21
struct Match
{
    size_t feature_index_1;
    size_t feature_index_2;
    double distance;
};

constexpr int DESCRIPTOR_BITS = 486; // include/opencalibration/types/feature_2d.hpp:11
constexpr int DESCRIPTOR_WORDS = 8;  // sizeof(std::bitset<486>) == 64 on libstdc++/x86-64

enum ModelKind
{
    MODEL_HOMOGRAPHY = 0,  // MINIMUM_POINTS 4, threshold 0.005 (homography_model.hpp:18,31)
    MODEL_ESSENTIAL = 1,   // MINIMUM_POINTS 5, threshold 0.01  (essential_matrix_model.hpp:19,31)
    MODEL_FUNDAMENTAL = 2, // MINIMUM_POINTS 8, threshold 0.01  (fundamental_matrix_model.hpp:19,29)
};
int minimum_points(int kind);
double default_threshold(int kind);

// A model = 3x3 matrix (column-major, like Eigen::Matrix3d) + its inverse for the homography.
struct Model
{
    int kind = MODEL_HOMOGRAPHY;
    double thr = 0.005;
    double M[9];    // homography / essential_matrix / fundamental_matrix, column-major
    double Minv[9]; // homography_inverse (H only)
    Model();
    explicit Model(int kind_);
};

// ---- src/match --------------------------------------------------------------------------------
// Inner loop of match_features_subset (src/match/match_features.cpp:71-98) on packed rows:
// best_k = POSITION of the first minimum, best_d/second_d = integer Hamming, 0xFFFF = +inf.
void match_top2(const uint64_t *q, size_t n1, const uint64_t *c, size_t n2, uint32_t *best_k, uint16_t *best_d,
                uint16_t *second_d);
// Column-wise top-1 (cross-check extension, not in the reference): for candidate j the first query
// position with the minimum distance; 0xFFFFFFFF when n1 == 0.
void match_col_best(const uint64_t *q, size_t n1, const uint64_t *c, size_t n2, uint32_t *col_best_q);
// src/match/match_features.cpp:54-103 verbatim semantics (gather, top-2, double ratio test, std::sort).
std::vector<Match> match_features_subset(const uint64_t *desc1, const uint64_t *desc2, const size_t *idx1, size_t n1,
                                         const size_t *idx2, size_t n2);
// src/match/match_features.cpp:8-52.
std::vector<size_t> spatially_subsample_feature_indices(const double *xy, const float *strength, size_t n_features,
                                                        double spacing_pixels, size_t count);

// ---- src/dense: guided matcher -----------------------------------------------------------------
// Inner loop of densifyMesh (src/dense/dense_stereo.cpp:251-273) on packed rows and CSR candidate lists: list l
// compares q[list_query[l]] with c[list_candidates[list_begin[l] .. list_begin[l+1])] in list order (the order of
// `nearby`, :244-246). best_pos = POSITION in the list of the first minimum (0 for an empty list, :253), distances in
// double exactly as descriptor_distance computes them (:56-59); +inf when absent.
void match_lists_top2(const uint64_t *q, const uint64_t *c, const uint32_t *list_query, const uint64_t *list_begin,
                      const uint32_t *list_candidates, size_t n_lists, uint32_t *best_pos, double *best_dist,
                      double *second_dist);
// The acceptance rule that follows (:275-276): >= 2 candidates: best < 0.85 * second; otherwise best < 0.35.
// (The reference never reaches it with an empty list, :248-249; an empty list is rejected here.)
bool guided_good_match(size_t list_length, double best_dist, double second_dist);

// ---- src/model_inliers: residuals ----------------------------------------------------------------
double error(const Model &m, const Corr &c);                                          // *_model.cpp ::error
double evaluate(const Model &m, const Corr *c, size_t n, std::vector<bool> &inliers); // *_model.cpp ::evaluate

// ---- src/model_inliers: fits ------------------------------------------------------------------
void fit(Model &m, const Corr *c, const size_t *sample);                               // ::fit
void fit_inliers(Model &m, const Corr *c, size_t n, const std::vector<bool> &inliers); // ::fitInliers
bool check_sample_degeneracy_h(const Corr *c, const size_t *sample);                   // homography_model.cpp:120-136
void check_degeneracy_f(Model &m, const Corr *c, size_t n,
                        std::vector<bool> &inliers); // fundamental_matrix_model.cpp:123-215

// ---- src/model_inliers/ransac.cpp -------------------------------------------------------------
struct RansacTrace
{
    size_t iterations = 0;   // loop iterations executed (value of i at exit)
    size_t improvements = 0; // times score > best_score
    size_t rejected = 0;     // SPRT rejections
    size_t degenerate = 0;   // checkSampleDegeneracy skips
};
// ransac.cpp:54-257. Returns evaluate(best)/N.
double ransac(const Corr *c, size_t n, Model &model, std::vector<bool> &inliers, RansacTrace *trace = nullptr);

// The hypothesis stream of ransac.cpp:98-171 with termination disabled: eval_order (after the
// shuffle) and the first `count` samples (MINIMUM_POINTS indices each). The stream depends only on
// n, the model kind and the quality ordering.
void hypothesis_stream(const Corr *c, size_t n, int kind, size_t count, std::vector<size_t> &eval_order,
                       std::vector<size_t> &samples);

// Per-hypothesis MSAC scoring without early exit: score summed in `order` (nullptr = 0..n-1), inlier
// count and bitmask (bit i of word i/32 = correspondence index i). ransac.cpp:183-196 / ::evaluate.
void score_hypothesis(const Model &m, const Corr *c, size_t n, const size_t *order, double *score, uint32_t *count,
                      uint32_t *bits);

// ---- linear algebra restated from Eigen 3.4.0 (column-major) ----------------------------------
namespace la
{
// FullPivLU(A).solve(b) for a rows x cols column-major A; x has cols entries.
void fullpivlu_solve(const double *A, int rows, int cols, const double *b, double *x);
// Matrix3d::inverse() (cofactor formula).
void inverse3(const double *M, double *Minv);
// JacobiSVD of a square n x n column-major matrix: U, V (n x n column-major), singular values sorted
// descending. U or V may be nullptr.
void jacobi_svd_square(const double *A, int n, double *U, double *S, double *V);
// Right singular vectors (cols x cols, column-major) + singular values of a rows x cols matrix with
// rows >= cols (Householder-QR preconditioned Jacobi).
void jacobi_svd_tall_v(const double *A, int rows, int cols, double *S, double *V);
} // namespace la

} // namespace oc_oracle
