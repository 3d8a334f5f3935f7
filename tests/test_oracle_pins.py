"""CPU tests that pin the oracle: against the reference's own object code (oracle/_ref), against the reference's
KATs and test tolerances, and against the committed golden vectors (tests/golden, generated from oracle/_ref)."""
import numpy as np
import pytest

import oc_oracle as O
from opencalibration_b200 import synthetic


def m33(M18):
    return np.asarray(M18[:9]).reshape(3, 3).T


def model_error(M, gt):
    a, b = M / np.linalg.norm(M), gt / np.linalg.norm(gt)
    return min(np.linalg.norm(a - b), np.linalg.norm(a + b))


def precision_recall(inl, n_true):
    gt = np.arange(len(inl)) < n_true
    tp, fp, fn = (inl & gt).sum(), (inl & ~gt).sum(), (~inl & gt).sum()
    return tp / max(tp + fp, 1), tp / max(tp + fn, 1)


# ---- layout facts the device rows rely on ------------------------------------------------
def test_reference_struct_layout(reference):
    lay = reference.layout()
    assert lay == dict(sizeof_feature_2d=96, offsetof_descriptor=24, sizeof_feature_match=24,
                       sizeof_correspondence=56)


# ---- src/match -------------------------------------------------------------------------------------------------
def test_subsample_kat(oracle):
    # test/test_match.cpp:90-107: 5 hand-made features, spacing 20 -> exactly {1, 3, 4}
    xy = np.array([[100, 100], [110, 105], [200, 200], [205, 202], [300, 300]], float)
    st = np.array([0.5, 0.9, 0.3, 0.7, 0.4], np.float32)
    assert oracle.subsample(xy, st, 20.0).tolist() == [1, 3, 4]


def test_subsample_matches_reference_on_real_features(oracle, reference, config1):
    for tag in "ab":
        for spacing in (40.0, 60.0, 160.0):
            a = oracle.subsample(config1[f"{tag}_xy"], config1[f"{tag}_strength"], spacing)
            b = reference.subsample(config1[f"{tag}_xy"], config1[f"{tag}_strength"], spacing)
            assert np.array_equal(a, b)
    # count argument (link_stage.cpp:64 passes num_sparse_features)
    a = oracle.subsample(config1["a_xy"], config1["a_strength"], 40.0, count=3000)
    b = reference.subsample(config1["a_xy"], config1["a_strength"], 40.0, count=3000)
    assert np.array_equal(a, b) and a.max() < 3000


def test_subsample_properties(oracle, config1):
    # test/test_match.cpp:45-88: monotone in spacing; all pairwise distances > spacing
    xy, st = config1["a_xy"], config1["a_strength"]
    n40, n80, n160 = (len(oracle.subsample(xy, st, s)) for s in (40.0, 80.0, 160.0))
    assert len(xy) > n40 > n80 > n160 > 5
    idx = oracle.subsample(xy, st, 60.0)
    p = xy[idx]
    d2 = ((p[:, None, :] - p[None, :, :]) ** 2).sum(-1)
    d2[np.arange(len(p)), np.arange(len(p))] = np.inf
    assert d2.min() > 60.0 ** 2


def test_subsample_golden(oracle, config1, golden):
    assert np.array_equal(oracle.subsample(config1["a_xy"], config1["a_strength"], 40.0), golden["c1_idx_a"])
    assert np.array_equal(oracle.subsample(config1["b_xy"], config1["b_strength"], 40.0), golden["c1_idx_b"])


def test_match_config1_golden(oracle, config1, golden):
    # configs[0]: the repo pair through test_match's flow; golden = the reference's own match_features.cpp
    i1, i2, d = oracle.match_features_subset(config1["a_desc"], config1["b_desc"], golden["c1_idx_a"],
                                             golden["c1_idx_b"])
    assert len(i1) == 500
    assert np.array_equal(i1, golden["c1_m1"]) and np.array_equal(i2, golden["c1_m2"])
    assert np.array_equal(d, golden["c1_dist"])
    # test/test_match.cpp:26-42: > 5 matches, original indices that are members of the passed subsets
    assert set(i1) <= set(golden["c1_idx_a"].tolist()) and set(i2) <= set(golden["c1_idx_b"].tolist())
    assert np.all(np.diff(d) <= 0)  # sorted by distance descending


def test_match_small_golden_with_ties(oracle, golden):
    i1, i2, d = oracle.match_features_subset(golden["c2s_a"], golden["c2s_b"], golden["c2s_i1"], golden["c2s_i2"])
    assert np.array_equal(i1, golden["c2s_m1"]) and np.array_equal(i2, golden["c2s_m2"])
    assert np.array_equal(d, golden["c2s_dist"])


@pytest.mark.parametrize("n1,n2", [(0, 0), (0, 5), (5, 0), (1, 1), (7, 1), (1, 9), (33, 65), (257, 300)])
def test_match_port_equals_reference_code(oracle, reference, n1, n2):
    a, b = synthetic.config2_pair(max(n1, 1) + 20, max(n2, 1) + 20, seed=100 + n1 + n2)
    if n2 > 3:
        b[2] = b[0]  # duplicates: equal distances exercise the second-best tie rule (match_features.cpp:88-91)
    rng = np.random.default_rng(n1 * 31 + n2)
    i1 = rng.permutation(len(a))[:n1]
    i2 = rng.permutation(len(b))[:n2]
    ra = oracle.match_features_subset(a, b, i1, i2)
    rb = reference.match_features_subset(a, b, i1, i2)
    assert all(np.array_equal(x, y) for x, y in zip(ra, rb))


def test_match_top2_semantics(oracle):
    # A3/A4 of SURVEY appendix A on hand-made rows
    z = np.zeros((1, 8), np.uint64)
    c = np.zeros((4, 8), np.uint64)
    c[0, 0] = 0b111      # d = 3
    c[1, 0] = 0b1        # d = 1  <- best (first minimum)
    c[2, 0] = 0b10       # d = 1  <- equal later distance becomes second best
    c[3, 0] = 0b1111     # d = 4
    bk, bd, sd = oracle.match_top2(z, c)
    assert (bk[0], bd[0], sd[0]) == (1, 1, 1)
    bk, bd, sd = oracle.match_top2(z, c[:1])
    assert (bk[0], bd[0], sd[0]) == (0, 3, 0xFFFF)      # one candidate: second = +inf, ratio test passes
    bk, bd, sd = oracle.match_top2(z, c[:0])
    assert (bk[0], bd[0], sd[0]) == (0, 0xFFFF, 0xFFFF)  # no candidate: best = +inf, index 0
    i1, i2, d = oracle.match_features_subset(z, c, [0], [0, 1, 2, 3])
    assert len(i1) == 0                                   # best == second -> never emitted
    i1, i2, d = oracle.match_features_subset(z, c, [0], [3, 0])
    assert (i1.tolist(), i2.tolist()) == ([0], [0]) and d[0] == 3 * (1.0 / 486)  # 3/486 < 0.8 * 4/486


# ---- linear algebra restated from Eigen ----------------------------------------------------------------------------
def test_linalg_against_numpy(oracle):
    rng = np.random.default_rng(0)
    for _ in range(5):
        A, b = rng.standard_normal((9, 9)), rng.standard_normal(9)
        assert np.allclose(oracle.fullpivlu_solve(A, b), np.linalg.solve(A, b), rtol=1e-9, atol=1e-11)
        M = rng.standard_normal((3, 3))
        assert np.allclose(oracle.inverse3(M), np.linalg.inv(M), rtol=1e-10, atol=1e-12)
        U, S, V = oracle.jacobi_svd_square(A)
        assert np.allclose(U @ np.diag(S) @ V.T, A, atol=1e-12)
        assert np.allclose(S, np.linalg.svd(A)[1], atol=1e-12) and np.all(np.diff(S) <= 0)
        T = rng.standard_normal((30, 3))
        S3, V3 = oracle.jacobi_svd_tall_v(T)
        assert np.allclose(S3, np.linalg.svd(T)[1], atol=1e-12)
        assert np.allclose(np.abs(V3), np.abs(np.linalg.svd(T)[2].T), atol=1e-9)
    # consistent over-determined system: FullPivLU.solve returns the exact solution (homography_model.cpp:81)
    x = rng.standard_normal(9)
    T = rng.standard_normal((21, 9))
    assert np.allclose(oracle.fullpivlu_solve(T, T @ x), x, atol=1e-9)
    # rank-deficient: zero pivots are dropped, the returned vector still solves the system
    A = rng.standard_normal((9, 9))
    A[:, 8] = A[:, 0]
    b = A @ rng.standard_normal(9)
    assert np.allclose(A @ oracle.fullpivlu_solve(A, b), b, atol=1e-9)


# ---- the reference's ransac unit tests, restated (test/test_ransac_unit.cpp) ---------------------------------------
SQUARE = [(1, 2, 1), (2, 2, 1), (2, 1, 1), (1, 1, 1)]


def corr_from(points1, points2=None, normalize=False):
    points2 = points1 if points2 is None else points2
    c = np.zeros((len(points1), 7))
    c[:, 0:3], c[:, 3:6] = np.asarray(points1, float), np.asarray(points2, float)
    if normalize:
        c[:, 0:3] /= np.linalg.norm(c[:, 0:3], axis=1, keepdims=True)
        c[:, 3:6] /= np.linalg.norm(c[:, 3:6], axis=1, keepdims=True)
    return c


@pytest.mark.parametrize("kind", [O.KIND_H, O.KIND_E, O.KIND_F])
def test_ransac_empty(oracle, kind):
    # *.ransac_compiles: empty input -> score 0, no inliers (test_ransac_unit.cpp:7-20,54-67,263-273)
    s, M, inl, _ = oracle.ransac(kind, np.zeros((0, 7)))
    assert s == 0 and len(inl) == 0
    s, M, inl, _ = oracle.ransac(kind, corr_from(SQUARE[:3]))  # fewer than MINIMUM_POINTS
    assert s == 0 and len(inl) == 3 and not inl.any()


def test_ransac_homography_fits_identity(oracle):
    # test_ransac_unit.cpp:22-52
    s, M, inl, _ = oracle.ransac(O.KIND_H, corr_from(SQUARE))
    assert s == pytest.approx(1.0, abs=4e-16 * 4) and inl.sum() == 4
    assert np.linalg.norm(m33(M) - np.eye(3)) < 1e-14


def test_ransac_fundamental_fits_identity(oracle):
    # test_ransac_unit.cpp:69-108
    pts = SQUARE + [(1, 2, 3), (2, 2, 2), (2, 1, 3), (1, 1, 2)]
    c = corr_from(pts, normalize=True)
    s, M, inl, _ = oracle.ransac(O.KIND_F, c)
    assert s == pytest.approx(1.0, abs=1e-15) and inl.sum() == 8
    assert abs(np.linalg.norm(m33(M)) - 1) < 1e-14
    assert sum(oracle.error(O.KIND_F, M, ci) for ci in c) < 1e-10


def perspective(v, R, T):
    ray = R.T @ (np.asarray(v, float) - (np.array([0, 0, 10.0]) + T))
    return np.array([ray[0] / ray[2] * 600, ray[1] / ray[2] * 600, 1.0])


@pytest.mark.parametrize("rot", [0.0, -np.pi / 2])
@pytest.mark.parametrize("T", [(0, 0, 0), (1, 0, 0), (1, -1, 0), (-1, 1, 0), (-1, -1, 0)])
def test_ransac_homography_rotation_translation(oracle, rot, T):
    # ransac_p.homography_rotation_translation (test_ransac_unit.cpp:114-176) without the decompose part
    down = np.diag([1.0, -1.0, -1.0])
    Rz = np.array([[np.cos(rot), -np.sin(rot), 0], [np.sin(rot), np.cos(rot), 0], [0, 0, 1]])
    T = np.array(T, float)
    p1, p2 = [], []
    for i in range(2):
        for j in range(2):
            p = (-1 if i > 0 else 1, -1 if j > 0 else 1, 0)
            p2.append(perspective(p, down, np.zeros(3)))
            p1.append(perspective(p, Rz @ down, T))
    c = corr_from(p1, p2)
    s, M, inl, _ = oracle.ransac(O.KIND_H, c)
    assert s == pytest.approx(1.0, abs=1e-15) and inl.sum() == 4
    H = m33(M)
    for a, b in zip(p1, p2):
        q = H @ a
        assert np.allclose(q / q[2], b, atol=1e-7)


def subset_scene():
    good = [(1, 2, 1), (2, 2, 1), (2, 1, 1), (1, 1, 1), (1.5, 1.5, 1), (1.2, 1.8, 1), (1.8, 1.2, 1), (1.3, 1.7, 1),
            (1, 2, 3), (2, 2, 2)]
    bad1 = [(100, 200, 1), (150, 250, 1), (120, 220, 1), (130, 230, 1)]
    bad2 = [(200, 100, 1), (250, 150, 1), (220, 120, 1), (230, 130, 1)]
    order = [0, "b0", 1, "b1", 2, 3, 4, "b2", 5, "b3", 6, 7, 8, 9]
    p1, p2, inl = [], [], []
    for o in order:
        if isinstance(o, str):
            k = int(o[1])
            p1.append(bad1[k]); p2.append(bad2[k]); inl.append(False)
        else:
            p1.append(good[o]); p2.append(good[o]); inl.append(True)
    return corr_from(p1, p2, normalize=True), np.array(inl)


def test_fundamental_fit_inliers_uses_subset(oracle):
    # test_ransac_unit.cpp:178-233
    c, inl = subset_scene()
    M = oracle.fit_inliers(O.KIND_F, np.full(18, np.nan), c, inl)
    err = np.array([abs(oracle.error(O.KIND_F, M, ci)) for ci in c])
    assert err[inl].mean() < 0.01 and err[~inl].mean() > 2 * err[inl].mean()


def test_fundamental_evaluate_uses_absolute_error(oracle):
    # test_ransac_unit.cpp:235-261
    pts = SQUARE + [(1.5, 1.5, 1), (1.2, 1.8, 1), (1.8, 1.2, 1), (1.3, 1.7, 1), (1, 2, 3), (2, 2, 2)]
    s, M, inl, _ = oracle.ransac(O.KIND_F, corr_from(pts, normalize=True))
    assert s > 0.7 and inl.sum() >= 8


def test_essential_fits_identity(oracle):
    # test_ransac_unit.cpp:275-298
    pts = SQUARE + [(1, 2, 3), (2, 2, 2)]
    s, M, inl, _ = oracle.ransac(O.KIND_E, corr_from(pts, normalize=True))
    assert s >= 0.16 and len(inl) == 6 and inl.sum() >= 1


# ---- test/test_ransac_benchmark.cpp floors ---------------------------------------------------------------------------
@pytest.mark.parametrize("n_in,n_out,pmin,rmin,emax", [(200, 0, 0.99, 0.99, 1e-6), (140, 60, 0.90, 0.85, None),
                                                      (80, 120, 0.80, 0.70, None), (40, 160, 0.70, 0.60, None)])
def test_benchmark_homography(oracle, n_in, n_out, pmin, rmin, emax):
    corr, gt = oracle.scene_homography(n_in, n_out, 42)
    s, M, inl, _ = oracle.ransac(O.KIND_H, corr)
    p, r = precision_recall(inl, n_in)
    assert p >= pmin and r >= rmin
    if emax:
        assert model_error(m33(M), gt) < emax


def test_benchmark_homography_near_degenerate(oracle):
    corr, gt = oracle.scene_homography_near_degenerate()
    s, M, inl, _ = oracle.ransac(O.KIND_H, corr)
    p, r = precision_recall(inl, 100)
    assert p >= 0.95 and r >= 0.95 and model_error(m33(M), gt) < 1e-6


@pytest.mark.parametrize("n_in,n_out,planar,pmin,rmin", [(200, 0, 0.0, 0.95, 0.80), (140, 60, 0.0, 0.85, 0.70),
                                                        (200, 0, 0.8, 0.95, 0.95)])
def test_benchmark_fundamental(oracle, n_in, n_out, planar, pmin, rmin):
    corr, gt = oracle.scene_fundamental(n_in, n_out, planar, 42)
    s, M, inl, _ = oracle.ransac(O.KIND_F, corr)
    p, r = precision_recall(inl, n_in)
    assert p >= pmin and r >= rmin


# ---- the RANSAC driver restatement equals the reference's own ransac.cpp -----------------------------------------------
def test_ransac_driver_equals_reference_code(oracle, reference):
    scenes = [(O.KIND_H, oracle.scene_homography(140, 60)[0]), (O.KIND_H, oracle.scene_homography(40, 160)[0]),
              (O.KIND_F, oracle.scene_fundamental(140, 60)[0]), (O.KIND_E, oracle.scene_fundamental(100, 20)[0]),
              (O.KIND_F, oracle.scene_fundamental(200, 0, 0.8)[0])]
    hq = oracle.scene_homography(300, 200, 7)[0]
    hq[:, 6] = np.random.default_rng(1).uniform(0.05, 0.4, len(hq))  # PROSAC branch (ransac.cpp:83-90,130-154)
    scenes.append((O.KIND_H, hq))
    for kind, corr in scenes:
        s1, M1, i1, _ = oracle.ransac(kind, corr)
        s2, M2, i2 = reference.ransac(kind, corr)
        assert s1 == s2 and np.array_equal(M1[:9], M2[:9]) and np.array_equal(i1, i2)


def test_ransac_golden(oracle, golden):
    names = sorted(k[3:-5] for k in golden.files if k.startswith("rs_") and k.endswith("_corr"))
    assert len(names) >= 8
    for name in names:
        kind = int(golden[f"rs_{name}_kind"]) if f"rs_{name}_kind" in golden.files else O.KIND_H
        s, M, inl, _ = oracle.ransac(kind, golden[f"rs_{name}_corr"])
        assert s == float(golden[f"rs_{name}_score"]), name
        assert np.array_equal(M[:9], golden[f"rs_{name}_M"][:9]), name
        assert np.array_equal(inl, golden[f"rs_{name}_inl"]), name


def test_config1_ransac_functional(oracle, golden):
    # test/test_ransac_functional.cpp:12-44: real pair, f = 5000 px pinhole, score > 0.20
    s, M, inl, tr = oracle.ransac(O.KIND_H, golden["rs_c1_corr"])
    assert s > 0.20 and inl.sum() > 6


def test_hypothesis_stream_is_score_independent(oracle):
    # the stream depends only on n / model / quality order
    corr, _ = oracle.scene_homography(140, 60)
    eo1, sm1 = oracle.hypothesis_stream(O.KIND_H, corr, 64)
    corr2 = corr.copy()
    corr2[:, 0:6] = np.random.default_rng(0).standard_normal((len(corr), 6))
    eo2, sm2 = oracle.hypothesis_stream(O.KIND_H, corr2, 64)
    assert np.array_equal(eo1, eo2) and np.array_equal(sm1, sm2)
    assert sorted(eo1.tolist()) == list(range(len(corr)))
    assert all(len(set(s.tolist())) == 4 for s in sm1)


def test_score_hypothesis_orders(oracle):
    corr, gt = oracle.scene_homography(300, 100)
    eo, sm = oracle.hypothesis_stream(O.KIND_H, corr, 8)
    models = np.stack([oracle.fit(O.KIND_H, corr, s) for s in sm])
    s_nat, c_nat, b_nat = oracle.score_hypotheses(O.KIND_H, models, corr, None, 0.005)
    s_eo, c_eo, b_eo = oracle.score_hypotheses(O.KIND_H, models, corr, eo, 0.005)
    assert np.array_equal(c_nat, c_eo) and np.array_equal(b_nat, b_eo)
    assert np.allclose(s_nat, s_eo, rtol=1e-12)
    for i in range(len(models)):
        s, inl = oracle.evaluate(O.KIND_H, models[i], corr)
        assert s == s_nat[i] and inl.sum() == c_nat[i]
