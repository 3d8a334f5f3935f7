/* ocb.h -- C ABI of the B200-native matching + RANSAC-scoring path (libocb.so).
 *
 * The reference (jkflying/opencalibration) has no FFI for this path: it is two static C++ libraries,
 * oc_match (src/match/CMakeLists.txt:1-6) and oc_model_inliers (src/model_inliers/CMakeLists.txt:1-8),
 * called from src/pipeline/link_stage.cpp:75-112. The C++ adapters in opencalibration_b200/host/ keep those
 * C++ signatures (include/opencalibration/match/match_features.hpp:10-16,
 * include/opencalibration/model_inliers/ransac.hpp:15-20) and reach the GPU only through the entry points
 * below. Each entry point names the reference code it replaces. Paths are relative to the reference root.
 *
 * Conventions
 *   - plain pointers and sizes; the caller owns every buffer; nothing is retained after a call returns
 *     (except descriptor sets registered with ocb_register_descriptors, which are copied to the device);
 *   - every function returning int returns 0 on success or a negative code (-(cudaError_t) for CUDA
 *     failures, OCB_E_* below otherwise); ocb_last_error() gives the message for the calling thread;
 *   - there is no CPU fallback: without a usable CUDA device every compute entry point fails;
 *   - host-buffer entry points are thread-safe and blocking (the reference calls this path concurrently
 *     from OpenMP workers, src/pipeline/pipeline.cpp:42-49); each calling thread gets its own stream and
 *     staging buffers on the device selected by ocb_set_device (default: the device of the process's first
 *     ocb_init, else device 0);
 *   - *_device entry points take device pointers (16-byte aligned) and a cudaStream_t passed as void*,
 *     enqueue asynchronously and do not synchronise.
 *
 * Descriptor rows: one row = the 64-byte memory image of std::bitset<486>
 * (include/opencalibration/types/feature_2d.hpp:11,15): 8 little-endian uint64 words, bits 486..511 zero.
 */
#ifndef OCB_H
#define OCB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C"
{
#endif

#define OCB_DESCRIPTOR_BITS 486
#define OCB_ROW_BYTES 64
#define OCB_ROW_WORDS 8
#define OCB_DIST_INF 0xFFFFu /* "+infinity": no (second) candidate seen */
#define OCB_NO_INDEX 0xFFFFFFFFu

#define OCB_E_INVALID (-10001)   /* bad argument (null pointer, misaligned device pointer, size overflow) */
#define OCB_E_NOT_FOUND (-10002) /* unknown descriptor-set id */
#define OCB_E_NO_DEVICE (-10003) /* no CUDA device / driver */

    /* Result of the top-2 search for one query row. Replaces the per-query state of the inner loop at
     * src/match/match_features.cpp:74-93: best_k is the POSITION (in the candidate array as passed) of the
     * first minimum, best_d / second_d are integer Hamming distances in [0,486]; second_d counts
     * multiplicity (a later candidate at the same distance as the best becomes the second best, :88-91).
     * n2 == 0 gives {0, OCB_DIST_INF, OCB_DIST_INF}; n2 == 1 gives second_d == OCB_DIST_INF. */
    typedef struct ocb_top2
    {
        uint32_t best_k;
        uint16_t best_d;
        uint16_t second_d;
    } ocb_top2;

    /* One directed image pair of a batched submission (unit of data parallelism of
     * src/pipeline/link_stage.cpp:75-112): query set -> candidate set, both registered beforehand. */
    typedef struct ocb_pair
    {
        uint64_t query_set;
        uint64_t candidate_set;
    } ocb_pair;

    /* One match that passed the ratio test, as the device emits it (ocb_match_pairs_ratio): the query POSITION in
     * the query set, the candidate POSITION in the candidate set and the integer Hamming distance; the reference's
     * feature_match (include/opencalibration/types/feature_match.hpp:10-21) is
     * {indices_1[query_k], indices_2[best_k], best_d * (1.0 / 486)}. */
    typedef struct ocb_match
    {
        uint32_t query_k;
        uint32_t best_k;
        uint32_t best_d;
    } ocb_match;

    /* The members of DifferentiableCameraModel<double> that image_to_3d reads
     * (include/opencalibration/types/camera_model.hpp:22-37). */
    typedef struct ocb_camera
    {
        double focal_length_pixels;
        double principal_point[2];
        double radial_distortion[3];
        double tangential_distortion[2];
        int32_t projection_planar; /* 1 = ProjectionType::PLANAR, 0 = UNKNOWN (the ray stays unset: NaN here) */
        int32_t reserved;
    } ocb_camera;

    enum ocb_model_kind
    {
        OCB_MODEL_HOMOGRAPHY = 0, /* include/opencalibration/model_inliers/homography_model.hpp:14-34 */
        OCB_MODEL_ESSENTIAL = 1,  /* .../essential_matrix_model.hpp:15-33 */
        OCB_MODEL_FUNDAMENTAL = 2 /* .../fundamental_matrix_model.hpp:15-31 */
    };

    /* ---- lifecycle ------------------------------------------------------------------------------ */
    int ocb_device_count(void);
    /* idempotent; selects `device` for the calling thread. The first successful call of the process also makes
     * `device` the default of every thread that never selects one itself (the reference's OpenMP workers,
     * src/pipeline/pipeline.cpp:42-49, in a one-process-per-GPU job). */
    int ocb_init(int device);
    int ocb_set_device(int device); /* device used by the calling thread's host-buffer calls */
    int ocb_current_device(void);   /* the device the calling thread's next host-buffer call will use */
    void ocb_shutdown(void);        /* frees cached staging buffers and registered descriptor sets */
    const char *ocb_last_error(void);
    const char *ocb_version(void);
    /* Number of kernels this library has launched since load (all threads). */
    uint64_t ocb_kernel_launches(void);
    /* Tuning knobs for experiments; results never depend on them (every setting is parity-tested). Unknown key:
     * OCB_E_INVALID.
     *   "k1_variant"       0 = default; instantiations of the Hamming kernel (queries per thread, adders, form)
     *   "k1_items_per_sm"  work items per SM when one pair's candidate axis is split (default 32)
     *   "k1_update"        0 = by run length, 1 = compare + vote + skip, 2 = branch-free two-smallest
     *   "k1_bf_rows"       runs shorter than this use the branch-free update when k1_update == 0
     *   "k1_engine"        single-pair search: 0 = by size (tensor cores from 512 x 512 rows up), 1 = integer pipes
     *                      (K1: XOR + POPC), 2 = tensor cores (K1T: exact s8 contraction, tcgen05.mma)
     *   "k1t_variant"      form of K1T's search kernel: 0 = default (4), 1 = first form, 2 / 3 = second form with one /
     *                      two query tiles per CTA (operands from shared memory), 4 = third form (query operand in
     *                      tensor memory)
     *   "k2_variant"       0 = by size, 1 = one hypothesis group per CTA, 2 = four groups per CTA in lock-step
     *   "k2_hg"            hypotheses per group (1..8); 0 = balance the SMs */
    int ocb_set_option(const char *key, int64_t value);
    /* Per-thread: 1 = the calling thread's host-buffer calls SLEEP while they wait for the device (an event with
     * blocking synchronisation) instead of spinning on the stream; 0 (default) = spin, the lowest latency for a lone
     * caller. A caller that runs more host threads than it has cores (the batched LinkStage runner on a box with four
     * cores per GPU) wants the core for its other threads; results never depend on it. */
    int ocb_set_thread_blocking_sync(int on);
    int64_t ocb_get_option(const char *key);

    /* ---- K1: Hamming top-2 --------------------------------------------------------------------------
     * Replaces the loop nest of match_features_subset, src/match/match_features.cpp:71-93, on packed rows
     * (the reference packs set_2 the same way, :62-66). q: [n1][8] uint64 query rows, c: [n2][8] candidate
     * rows, out: [n1]. col_best_q (nullable, [n2]): cross-check extension that the reference does not have --
     * for every candidate the position of the first query at minimum distance (OCB_NO_INDEX if n1 == 0),
     * computed by a second pass with the roles swapped. The double-precision ratio test (:94) and the
     * std::sort (:100-101) stay in the C++ adapter so their results are the reference's bit for bit. */
    int ocb_match_top2(const uint64_t *q, size_t n1, const uint64_t *c, size_t n2, ocb_top2 *out,
                       uint32_t *col_best_q);

    /* Same search on rows that are NOT contiguous: row k of the query side is the 64 bytes at
     * rows1 + idx1[k] * stride1 (idx1 == NULL: k * stride1), likewise the candidate side. This is the access
     * pattern of match_features_subset itself -- set[indices[k]].descriptor inside std::vector<feature_2d>
     * (sizeof 96, descriptor at +24; src/match/match_features.cpp:62-66,71) -- so the adapter passes the vectors
     * as they are and the rows are gathered once, straight into page-locked staging. best_k / col_best_q are
     * POSITIONS k in idx2 / idx1. */
    int ocb_match_top2_strided(const void *rows1, size_t stride1, const size_t *idx1, size_t n1, const void *rows2,
                               size_t stride2, const size_t *idx2, size_t n2, ocb_top2 *out, uint32_t *col_best_q);

    /* Device-resident variant. d_q, d_c, d_out, d_col_best_q, d_workspace are device pointers;
     * workspace_bytes >= ocb_match_top2_workspace_bytes(n1, n2, d_col_best_q != NULL). */
    size_t ocb_match_top2_workspace_bytes(size_t n1, size_t n2, int with_col_best);
    int ocb_match_top2_device(const void *d_q, size_t n1, const void *d_c, size_t n2, void *d_out,
                              void *d_col_best_q, void *d_workspace, size_t workspace_bytes, void *stream);

    /* ---- descriptor residency + batched pairs (LinkStage granularity) ---------------------------------
     * Registers (copies to the calling thread's device) the packed rows of one image's subsampled features,
     * i.e. set[indices[k]] for the indices spatially_subsample_feature_indices returned
     * (src/pipeline/link_stage.cpp:63-65,80-81). Re-registering an id replaces it. */
    int ocb_register_descriptors(uint64_t set_id, const uint64_t *rows, size_t n);
    /* A set must not be unregistered or re-registered while an ocb_match_pairs call that names it is in flight on
     * another thread (the storage goes back to the device's memory pool without waiting for other threads' streams). */
    int ocb_unregister_descriptors(uint64_t set_id);
    /* Registers many sets with ONE device allocation and a pipelined gather -> page-locked staging -> device copy
     * (the per-image cudaMalloc + synchronous copy of ocb_register_descriptors dominates when a LinkStage batch
     * uploads hundreds of images). Row k of a set is the 64 bytes at rows + idx[k] * stride (idx == NULL: k * stride):
     * pass &features[0].descriptor, sizeof(feature_2d) and the subsample indices to upload straight from
     * std::vector<feature_2d>. The allocation is released when the last of its sets is unregistered / replaced.
     * A set_id may appear only once per batch (OCB_E_INVALID otherwise). */
    typedef struct ocb_set_source
    {
        uint64_t set_id;
        const void *rows;
        size_t stride;
        const size_t *idx;
        size_t n;
    } ocb_set_source;
    int ocb_register_descriptors_batch(const ocb_set_source *sources, size_t count);
    /* Like ocb_register_descriptors_batch, plus what the steps AFTER the match need on the device (K6): the keypoint
     * location of every row (feature_2d::location, two doubles at xy + idx[k] * xy_stride; xy == NULL: the set has no
     * keypoints and cannot be named by ocb_corr_bind_batch_matches) and the image's camera model. */
    typedef struct ocb_image_source
    {
        uint64_t set_id;
        const void *rows;
        size_t stride;
        const size_t *idx;
        size_t n;
        const void *xy;
        size_t xy_stride;
        ocb_camera camera;
    } ocb_image_source;
    int ocb_register_images_batch(const ocb_image_source *sources, size_t count);
    /* Page-locked host memory for callers that want results copied straight into their buffers (every host-buffer
     * entry point detects page-locked arguments and skips its staging copy). NULL on failure. */
    void *ocb_host_alloc(size_t bytes);
    void ocb_host_free(void *p);
    /* Matches n_pairs pairs in one submission. out receives the ocb_top2 records of pair p at
     * out[out_offsets[p] .. out_offsets[p] + n_query_rows(p)); out_offsets has n_pairs entries. */
    int ocb_match_pairs(const ocb_pair *pairs, size_t n_pairs, ocb_top2 *out, const uint64_t *out_offsets);

    /* ocb_match_pairs followed ON THE DEVICE by the ratio test of src/match/match_features.cpp:94 (best < 0.8 * second
     * in IEEE double on distance = count * (1.0 / 486), :79 -- the reference's comparison bit for bit) and an
     * order-preserving compaction (K5): the survivors of pair p, in query order (the emission order of :71-97), are
     * out[out_offsets[p] .. out_offsets[p + 1]); out_offsets has n_pairs + 1 entries and out_offsets[n_pairs] is the
     * total. Only the survivors cross PCIe (12 bytes each instead of 8 bytes per query row). out_capacity = records
     * `out` can hold; a larger total gives OCB_E_INVALID (the sum of the pairs' query rows always suffices). The
     * reference's std::sort (:100-101) is left to the caller. */
    int ocb_match_pairs_ratio(const ocb_pair *pairs, size_t n_pairs, ocb_match *out, size_t out_capacity,
                              uint64_t *out_offsets);
    /* ocb_match_pairs_ratio followed ON THE DEVICE by the reference's final std::sort (K7): pair p's survivors come
     * back in the order of src/match/match_features.cpp:100-101 -- descending distance, ties in the order libstdc++'s
     * (unstable) introsort leaves them, replayed step for step -- so out[out_offsets[p] + i] is match i of the
     * reference's result. quality_order (nullable, out_capacity entries): for the same pair,
     * quality_order[out_offsets[p] + j] = the index i of the match at rank j of the reference's PROSAC ordering
     * (src/model_inliers/ransac.cpp:83-90: std::sort of 0 .. n-1 by quality = distance, ascending). */
    int ocb_match_pairs_sorted(const ocb_pair *pairs, size_t n_pairs, ocb_match *out, size_t out_capacity,
                               uint64_t *out_offsets, uint32_t *quality_order);

    /* ---- K4: Hamming top-2 over per-query candidate lists (guided matcher of the dense stage) ---------------
     * Replaces the inner loop of densifyMesh, src/dense/dense_stereo.cpp:251-273: list l compares query row
     * q[list_query[l]] with the candidate rows c[list_candidates[list_begin[l] .. list_begin[l+1])] in list order
     * (the order of the KD-tree radius search result `nearby`, :244-246) under the update rule of
     * match_features.cpp:80-92. out[l].best_k is the POSITION within list l of the first minimum (the adapter maps it
     * to nearby[best_k].payload), best_d / second_d as in ocb_top2; an empty list gives {0, INF, INF}. list_begin has
     * n_lists + 1 non-decreasing entries; a list may hold at most OCB_MAX_LIST_LENGTH candidates. Indices out of
     * range give OCB_E_INVALID (checked on the host before anything is launched). The acceptance rule (:275-276)
     * stays in the adapter. */
#define OCB_MAX_LIST_LENGTH 0x3FFFFEu
    int ocb_match_lists(const uint64_t *q, size_t n_q_rows, const uint64_t *c, size_t n_c_rows,
                        const uint32_t *list_query, const uint64_t *list_begin, const uint32_t *list_candidates,
                        size_t n_lists, ocb_top2 *out);
    /* Device-resident variant: every pointer is a device pointer (rows 16-byte aligned), nothing is validated,
     * enqueues on `stream` and does not synchronise. */
    int ocb_match_lists_device(const void *d_q, const void *d_c, const void *d_list_query, const void *d_list_begin,
                               const void *d_list_candidates, size_t n_lists, void *d_out, void *stream);

    /* ---- K2/K3: hypotheses x correspondences MSAC scoring -------------------------------------------------
     * Replaces the score loop of ransac<Model>, src/model_inliers/ransac.cpp:183-196 (without the SPRT
     * early exit, which the host driver replays), and Model::evaluate
     * (src/model_inliers/homography_model.cpp:99-118, essential_matrix_model.cpp:91-110,
     * fundamental_matrix_model.cpp:89-108), for h hypotheses at once.
     *   models:  [h][18] doubles: the 3x3 matrix column-major (Eigen::Matrix3d storage) followed by, for the
     *            homography, homography_inverse; the last 9 are ignored for essential / fundamental.
     *   corr:    [n][7] doubles = std::vector<correspondence>::data()
     *            (include/opencalibration/types/correspondence.hpp:8-13).
     *   order:   nullable [n] permutation: the MSAC sum of each hypothesis is accumulated sequentially in
     *            this order (ransac.cpp's shuffled eval_order); NULL = index order (evaluate). An entry >= n gives
     *            OCB_E_INVALID (here, in ocb_corr_bind and in ocb_corr_bind_batch).
     *   score:   [h] sum of 1-(e/thr)^2 over inliers (e < thr, strict), IEEE double, each operation rounded
     *            individually (no FMA contraction), summed in `order`.
     *   count:   [h] number of inliers.
     *   inlier_bits: nullable [h][ceil(n/32)]: bit (i & 31) of word i/32 set iff correspondence i is an inlier. */
    int ocb_score_models(int kind, const double *models, size_t h, const double *corr, size_t n, double thr,
                         const uint32_t *order, double *score, uint32_t *count, uint32_t *inlier_bits);
    /* Per-correspondence residuals e[i] = Model::error(corr[i]) of ONE model (homography_model.cpp:89-97,
     * essential_matrix_model.cpp:112-123), index order; used by the host RANSAC driver to replay the SPRT
     * prefix test (ransac.cpp:197-202) exactly for the rare hypotheses that beat the best score. */
    int ocb_residuals(int kind, const double *model18, const double *corr, size_t n, double *e);

    /* One RANSAC run scores many batches of models (and single models: evaluate, residuals) against the SAME
     * correspondences (src/model_inliers/ransac.cpp:162-252). ocb_corr_bind uploads and prepares them once for the
     * calling thread (index order, and evaluation order when `order` is given); ocb_score_bound /
     * ocb_residuals_bound then behave exactly like ocb_score_models (order = in_order ? the bound order : NULL) and
     * ocb_residuals, without re-sending the correspondences. The binding is per thread and lasts until the next
     * ocb_corr_bind / ocb_corr_unbind on that thread. */
    int ocb_corr_bind(const double *corr, size_t n, const uint32_t *order);
    int ocb_corr_unbind(void);
    int ocb_score_bound(int kind, const double *models, size_t h, double thr, int in_order, double *score,
                        uint32_t *count, uint32_t *inlier_bits);
    int ocb_residuals_bound(int kind, const double *model18, double *e);
    /* Device-side minimal-sample fits against the bound correspondences (see OCB_REQ_FIT_SCORE_ORDERED below):
     * samples [h][4] -> models_out [h][18], degenerate [h]; with score != NULL the fitted models are also scored in the
     * bound evaluation order (score [h], count [h]) in the same submission. */
    int ocb_fit_score_bound(const uint32_t *samples, size_t h, double thr, double *models_out, uint8_t *degenerate,
                            double *score, uint32_t *count);
    /* One step of the local optimisation of src/model_inliers/ransac.cpp:224-245 against the bound correspondences:
     * homography_model::fitInliers (homography_model.cpp:52-87) ON THE DEVICE for the correspondences whose bit is set
     * in refit_bits [ceil(n/32)] (index order), then Model::evaluate of the refitted model -> model_out [18],
     * score [1], count [1], inlier_bits [ceil(n/32)]. Equals the host fitInliers + ocb_score_bound bit for bit. */
    int ocb_refit_evaluate_bound(const uint32_t *refit_bits, double thr, double *model_out, double *score,
                                 uint32_t *count, uint32_t *inlier_bits);
    /* Stand-alone form: fits h minimal samples of corr [n][7] (replaces homography_model::checkSampleDegeneracy +
     * homography_model::fit, src/model_inliers/homography_model.cpp:19-50,120-136, for h hypotheses at once).
     * Replaces the calling thread's ocb_corr_bind binding. */
    int ocb_fit_homography(const double *corr, size_t n, const uint32_t *samples, size_t h, double *models_out,
                           uint8_t *degenerate);

    /* Batched form for a whole submission of image pairs (the batched LinkStage runner advances the RANSAC runs of
     * all its pairs in lock step): ocb_corr_bind_batch makes `count` correspondence sets resident for the calling
     * thread with one copy; ocb_score_requests then serves any mix of requests against them with ONE kernel launch
     * and one copy each way:
     *   mode OCB_REQ_SCORE_ORDERED  h models scored in the set's evaluation order -> score[h], count[h]
     *                               (the score loop of src/model_inliers/ransac.cpp:183-196, no early exit)
     *   mode OCB_REQ_EVALUATE       h models in index order -> score[h], count[h], inlier_bits[h][ceil(n/32)]
     *                               (Model::evaluate, homography_model.cpp:99-118 and twins)
     *   mode OCB_REQ_RESIDUALS      one model -> residuals[n] (Model::error per correspondence)
     *   mode OCB_REQ_FIT_SCORE_ORDERED  (homography only) h minimal samples of 4 correspondence indices each are
     *                               checked (homography_model::checkSampleDegeneracy, homography_model.cpp:120-136),
     *                               fitted ON THE DEVICE (homography_model::fit, :19-50: 9x9 DLT system with the
     *                               h33 == 1 row, full-pivot LU solve, division by H(2,2), 3x3 inverse) and scored
     *                               like OCB_REQ_SCORE_ORDERED -> models_out[h][18], degenerate[h], score[h],
     *                               count[h]. This is ransac.cpp:164-196 for h iterations without the host in
     *                               the loop; the fitted models equal the host adapters' fit bit for bit.
     *   mode OCB_REQ_REFIT_EVALUATE (homography only, h == 1) one step of the local optimisation of
     *                               src/model_inliers/ransac.cpp:224-245 without the host in between: the model is
     *                               refitted ON THE DEVICE to the correspondences whose bit is set in refit_bits
     *                               (homography_model::fitInliers, homography_model.cpp:52-87: the (2m+1) x 9 DLT
     *                               system, full-pivot LU solve, division by H(2,2), 3x3 inverse) and then evaluated
     *                               like OCB_REQ_EVALUATE -> models_out[18], score[1], count[1], inlier_bits.
     * Results are bit-identical to ocb_score_models / ocb_residuals on the same inputs. */
    typedef struct ocb_corr_set
    {
        const double *corr;    /* [n][7] */
        size_t n;
        const uint32_t *order; /* nullable [n] */
    } ocb_corr_set;
    enum ocb_request_mode
    {
        OCB_REQ_SCORE_ORDERED = 0,
        OCB_REQ_EVALUATE = 1,
        OCB_REQ_RESIDUALS = 2,
        OCB_REQ_FIT_SCORE_ORDERED = 3,
        OCB_REQ_REFIT_EVALUATE = 4
    };
    typedef struct ocb_score_request
    {
        uint32_t set; /* index into the batch bound by ocb_corr_bind_batch */
        int32_t kind; /* enum ocb_model_kind */
        int32_t mode; /* enum ocb_request_mode */
        uint32_t h;   /* number of models (1 for OCB_REQ_RESIDUALS) */
        const double *models; /* [h][18] */
        double thr;
        double *score;
        uint32_t *count;
        uint32_t *inlier_bits;
        double *residuals;
        /* OCB_REQ_FIT_SCORE_ORDERED only (models is ignored): */
        const uint32_t *samples; /* [h][4] correspondence indices of the minimal samples */
        double *models_out;      /* [h][18] fitted models (NaN for a degenerate sample) */
        uint8_t *degenerate;     /* [h] 1 = the sample failed checkSampleDegeneracy and was not fitted */
        /* OCB_REQ_REFIT_EVALUATE only (models is ignored, models_out receives the refitted model): */
        const uint32_t *refit_bits; /* [ceil(n/32)] the inliers to refit to, index order */
    } ocb_score_request;
    int ocb_corr_bind_batch(const ocb_corr_set *sets, size_t count);
    /* distort_keypoints on the device (K6) feeding the same binding: replaces src/distort/distort_keypoints.cpp:48-61
     * (image_to_3d of both keypoints of every match, :62-103, quality = match distance) as LinkStage calls it
     * (src/pipeline/link_stage.cpp:87-88) for a whole batch of pairs, and leaves the correspondences bound for
     * ocb_score_requests exactly as ocb_corr_bind_batch would. matches[i] = {position in set_1, position in set_2,
     * integer distance} IN THE ORDER OF THE SORTED MATCH LIST (match_features.cpp:100-101): correspondence i belongs to
     * match i. Both sets must have been registered with keypoints (ocb_register_images_batch) on this thread's device.
     * corr_out (nullable) receives the [n][7] rows, bit-identical to the host distort_keypoints of the C++ mirror. */
    typedef struct ocb_match_set
    {
        uint64_t set_1, set_2;
        const ocb_match *matches; /* [n] */
        size_t n;
        const uint32_t *order;    /* nullable [n] evaluation order, as in ocb_corr_set */
        double *corr_out;         /* nullable [n][7] */
    } ocb_match_set;
    int ocb_corr_bind_batch_matches(const ocb_match_set *sets, size_t count);
    /* image_to_3d (src/distort/distort_keypoints.cpp:62-103) for n keypoints of one camera: xy [n][2] -> rays [n][3]. */
    int ocb_image_to_3d(const double *xy, size_t n, const ocb_camera *camera, double *rays);
    int ocb_score_requests(const ocb_score_request *requests, size_t count);

    /* Device-resident variant of ocb_score_models. d_corr4: [n][4] doubles (x1,y1,x2,y2) = measurement / z,
     * already in evaluation order; d_pos: nullable [n] uint32 correspondence index of each evaluation
     * position (for the bit mask); everything else as above but device pointers. */
    int ocb_score_models_device(int kind, const void *d_models, size_t h, const void *d_corr4, const void *d_pos,
                                size_t n, double thr, void *d_score, void *d_count, void *d_inlier_bits,
                                void *stream);
    /* corr [n][7] (device) -> corr4 [n][4] in `order` (device, nullable) + pos. */
    int ocb_prepare_correspondences_device(const void *d_corr7, const void *d_order, size_t n, void *d_corr4,
                                           void *d_pos, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* OCB_H */
