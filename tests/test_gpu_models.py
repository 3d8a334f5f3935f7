"""GPU parity tests for K2/K3 (MSAC scoring) and the RANSAC driver of the C++ mirror.
Bar: inlier sets, counts and MSAC scores bit-exact versus the CPU oracle (same canonical operation order, IEEE
double, sums in evaluation order); fitted models bit-exact versus the oracle's restatement and within 1e-9
(relative Frobenius) of what the reference's tests require."""
import numpy as np
import pytest

import oc_oracle as O
from opencalibration_b200 import synthetic

pytestmark = pytest.mark.gpu

THR = {0: 0.005, 1: 0.01, 2: 0.01}


def stream_models(oracle, kind, corr, count):
    eo, samples = oracle.hypothesis_stream(kind, corr, count)
    models = np.zeros((count, 18))
    for i, s in enumerate(samples):
        models[i] = np.nan_to_num(oracle.fit(kind, corr, s), nan=0.0)
    return eo, samples, models


@pytest.mark.parametrize("kind", [0, 1, 2])
@pytest.mark.parametrize("n", [1, 31, 224, 225, 1000, 3333])
def test_scores_bit_exact(gpu, oracle, kind, n):
    n_in = int(n * 0.7)
    corr = (oracle.scene_homography(n_in, n - n_in, 5)[0] if kind == 0 else
            oracle.scene_fundamental(n_in, n - n_in, 0.0, 5)[0])
    if n < O.MIN_POINTS[kind]:
        models = synthetic.random_models(kind, 9, seed=n)
        eo = np.random.default_rng(0).permutation(n)
    else:
        eo, _, models = stream_models(oracle, kind, corr, 19)
    for order in (None, eo):
        s, c, bits = gpu.score_models(kind, models, corr, THR[kind], order=order)
        so, co, bo = oracle.score_hypotheses(kind, models, corr, order=order, thr=THR[kind])
        assert np.array_equal(s, so)          # sequential double sum, bit for bit
        assert np.array_equal(c, co) and np.array_equal(bits, bo)


@pytest.mark.parametrize("spread", [4, 60, 700, 1000])
def test_shared_reciprocal_division_and_sqrt_equal_the_ieee_intrinsics(gpu, spread):
    """csrc/exact_math.cuh: K2 divides x and y by the same z (homography_model.cpp:91-96) through ONE refined
    reciprocal and takes its square root in line; both must return the bits of div.rn.f64 / sqrt.rn.f64."""
    counts = gpu.probe_exact_math(seed=spread * 7919, n=6_000_000, exponent_spread=spread)
    assert counts[0] > (5_000_000 if spread <= 60 else 100_000) and counts[2] > 1_000_000
    assert counts[1] == 0 and counts[3] == 0, counts


@pytest.mark.parametrize("kind", [0, 2])
@pytest.mark.parametrize("variant,hg", [(0, 0), (1, 0), (2, 0), (1, 8), (2, 5), (2, 4)])
def test_scores_bit_exact_with_extreme_operands(gpu, oracle, kind, variant, hg):
    """Operands outside the guarded range of the shared-reciprocal forms (zeros, denormals, huge values, exact fits)
    take the out-of-line IEEE path; every K2 variant / group size must still equal the oracle bit for bit."""
    n = 700
    corr = (oracle.scene_homography(500, 200, 11)[0] if kind == 0 else oracle.scene_fundamental(500, 200, 0.0, 11)[0])
    _, _, models = stream_models(oracle, kind, corr, 21)
    rng = np.random.default_rng(5)
    corr[0:40, 0:2] = 0.0                      # x1 = y1 = 0
    corr[40:60, 3:5] *= 1e-300                 # tiny second view
    corr[60:80, 0:2] *= 1e+250                 # huge first view
    corr[80:90, 0:6] = 0.0                     # z == 0 -> NaN
    models[3] = np.concatenate([np.eye(3).ravel(), np.eye(3).ravel()])   # exact fit of identical points: e == 0
    corr[100:140, 3:6] = corr[100:140, 0:3]
    models[4] *= 1e-200
    models[5] *= 1e+200
    models[6, :] = 0.0
    models[7, rng.integers(0, 9, 3)] = 0.0
    gpu.set_option("k2_variant", variant)
    gpu.set_option("k2_hg", hg)
    try:
        for thr in (THR[kind], 1e-300, 1e+300):
            for order in (None, rng.permutation(n)):
                s, c, bits = gpu.score_models(kind, models, corr, thr, order=order)
                so, co, bo = oracle.score_hypotheses(kind, models, corr, order=order, thr=thr)
                assert np.array_equal(s, so, equal_nan=True)
                assert np.array_equal(c, co) and np.array_equal(bits, bo)
    finally:
        gpu.set_option("k2_variant", 0)
        gpu.set_option("k2_hg", 0)


@pytest.mark.parametrize("kind", [0, 1])
@pytest.mark.parametrize("h,n", [(1200, 500), (4096, 2000), (601, 225)])
def test_lockstep_and_independent_forms_agree_with_the_oracle(gpu, oracle, kind, h, n):
    """Many hypotheses go out as 1024-thread CTAs holding four hypothesis groups in lock-step, few as one group per
    CTA (k2_variant 2 / 1; 0 picks by size). Scores, counts and masks must not depend on the form."""
    corr = (oracle.scene_homography(int(n * 0.7), n - int(n * 0.7), 9)[0] if kind == 0 else
            oracle.scene_fundamental(int(n * 0.7), n - int(n * 0.7), 0.0, 9)[0])
    _, _, base = stream_models(oracle, kind, corr, 32)
    models = np.tile(base, (h // 32 + 1, 1))[:h] * np.linspace(1.0, 1.5, h)[:, None]
    eo = np.random.default_rng(h).permutation(n)
    so, co, bo = oracle.score_hypotheses(kind, models, corr, order=eo, thr=THR[kind])
    try:
        for variant in (0, 1, 2):
            gpu.set_option("k2_variant", variant)
            s, c, bits = gpu.score_models(kind, models, corr, THR[kind], order=eo)
            assert np.array_equal(s, so) and np.array_equal(c, co) and np.array_equal(bits, bo), variant
    finally:
        gpu.set_option("k2_variant", 0)


def test_residuals_bit_exact_and_special_values(gpu, oracle):
    corr, _ = oracle.scene_homography(500, 200, 3)
    corr[10, 2] = 0.0       # measurement1.z == 0
    corr[11, 5] = np.inf    # measurement2.z == inf
    corr[12, 0] = np.nan
    corr[13, 0:6] *= -3.5   # homogeneous scale (incl. negative z) must not matter after m / m.z
    for kind in (0, 2):
        _, _, models = stream_models(oracle, kind, corr[20:], 3)
        for m in models:
            e = gpu.residuals(kind, m, corr)
            eo = np.array([oracle.error(kind, m, c) for c in corr])
            assert np.array_equal(e, eo, equal_nan=True)
            assert np.isnan(e[10]) and np.isnan(e[11]) and np.isnan(e[12])
        s, c, bits = gpu.score_models(kind, models, corr, THR[kind])
        so, co, bo = oracle.score_hypotheses(kind, models, corr, thr=THR[kind])
        assert np.array_equal(s, so) and np.array_equal(c, co) and np.array_equal(bits, bo)
        assert not (bits[:, 0] >> 10 & 1).any()  # NaN residuals are never inliers
    zero = np.zeros((1, 18))  # epipolar denominator < 1e-20 -> DBL_MAX -> no inliers
    s, c, bits = gpu.score_models(1, zero, corr, 0.01)
    assert s[0] == 0 and c[0] == 0
    assert gpu.residuals(1, zero[0], corr[:5]).tolist() == [np.finfo(np.float64).max] * 5


def test_threshold_is_strict(gpu, oracle):
    # e < t (ransac.cpp:189): move the threshold onto an observed residual and just above it
    corr, gt = oracle.scene_homography(50, 50, 8)
    _, _, models = stream_models(oracle, 0, corr, 4)
    e = gpu.residuals(0, models[0], corr)
    t = np.sort(e[np.isfinite(e)])[30]
    for thr in (t, np.nextafter(t, np.inf)):
        s, c, bits = gpu.score_models(0, models[:1], corr, thr)
        so, co, bo = oracle.score_hypotheses(0, models[:1], corr, thr=thr)
        assert np.array_equal(c, co) and np.array_equal(s, so) and np.array_equal(bits, bo)
    assert gpu.score_models(0, models[:1], corr, t)[1][0] + 1 == gpu.score_models(0, models[:1], corr,
                                                                                   np.nextafter(t, np.inf))[1][0]


def test_evaluation_order_entries_are_validated(gpu, oracle):
    # order[] indexes the correspondences on the device; an entry >= n must be an error, not a wild read
    corr, _ = oracle.scene_homography(40, 10, 1)
    M = np.zeros((2, 18))
    M[:, [0, 4, 8, 9, 13, 17]] = 1.0
    order = np.arange(len(corr), dtype=np.uint32)
    gpu.score_models(0, M, corr, 0.005, order=order)
    order[7] = len(corr)
    with pytest.raises(gpu.OcbError, match="order"):
        gpu.score_models(0, M, corr, 0.005, order=order)


def test_many_hypotheses_ragged_groups(gpu, oracle):
    corr, _ = oracle.scene_homography(700, 300, 11)
    eo, _, models = stream_models(oracle, 0, corr, 61)  # 61 = 7 full groups of 8 + 5
    s, c, bits = gpu.score_models(0, models, corr, 0.005, order=eo)
    so, co, bo = oracle.score_hypotheses(0, models, corr, order=eo, thr=0.005)
    assert np.array_equal(s, so) and np.array_equal(c, co) and np.array_equal(bits, bo)
    s2, c2, _ = gpu.score_models(0, models, corr, 0.005, order=eo, want_bits=False)
    assert np.array_equal(s2, s) and np.array_equal(c2, c)


def test_config3_full_size_every_hypothesis(gpu, oracle):
    # BASELINE configs[2]: 4k hypotheses x 20k correspondences from the reference's seeded stream; ALL 4096 scores,
    # counts and inlier bit masks (8.2e7 residuals) against the oracle, bit for bit
    corr, gt = oracle.scene_homography(14000, 6000, 42)
    eo, samples, models = stream_models(oracle, 0, corr, 4096)
    s, c, bits = gpu.score_models(0, models, corr, 0.005, order=eo)
    so, co, bo = oracle.score_hypotheses(0, models, corr, order=eo, thr=0.005)
    assert np.array_equal(s, so) and np.array_equal(c, co) and np.array_equal(bits, bo)
    assert (c > 10000).sum() > 500  # a good share of the stream's samples are all-inlier
    # size-independent properties: a sample drawn from inliers only scores >= 14000 * small, counts bound scores
    assert np.all(s <= c) and np.all(s >= 0) and c.max() >= 14000
    all_inlier_samples = np.all(samples < 14000, axis=1)
    assert c[all_inlier_samples].min() >= 13990
    # evaluation order changes neither counts nor (beyond rounding) scores
    s_nat, c_nat, _ = gpu.score_models(0, models[:64], corr, 0.005, want_bits=False)
    assert np.array_equal(c_nat, c[:64]) and np.allclose(s_nat, s[:64], rtol=1e-12)


def test_config3_epipolar_full_size_every_hypothesis(gpu, oracle):
    # the E/F residual on the same shape: SyntheticScene::fundamental(14000, 6000), threshold 0.01
    corr, gt = oracle.scene_fundamental(14000, 6000, 0.0, 42)
    eo, samples, models = stream_models(oracle, 2, corr, 4096)
    s, c, bits = gpu.score_models(2, models, corr, 0.01, order=eo)
    so, co, bo = oracle.score_hypotheses(2, models, corr, order=eo, thr=0.01)
    assert np.array_equal(s, so) and np.array_equal(c, co) and np.array_equal(bits, bo)


def test_evaluate_through_the_mirror(gpu, hostlib, oracle):
    for kind in (0, 1, 2):
        corr = (oracle.scene_homography(300, 100, 2)[0] if kind == 0 else oracle.scene_fundamental(300, 100, 0, 2)[0])
        _, _, models = stream_models(oracle, kind, corr, 5)
        for m in models:
            s, inl = hostlib.evaluate(kind, m, corr)
            so, io = oracle.evaluate(kind, m, corr)
            assert s == so and np.array_equal(inl, io)
    s, inl = hostlib.evaluate(0, models[0], np.zeros((0, 7)))
    assert s == 0 and len(inl) == 0


# ---- the RANSAC driver -------------------------------------------------------------------------------------------------
def m33(M18):
    return np.asarray(M18[:9]).reshape(3, 3).T


def check_ransac_equal(hostlib, oracle, kind, corr):
    s, M, inl, st = hostlib.ransac(kind, corr)
    so, Mo, io, tr = oracle.ransac(kind, corr)
    assert s == so
    assert np.array_equal(inl, io)
    n = 18 if kind == 0 else 9
    assert np.array_equal(M[:n], Mo[:n], equal_nan=True)
    assert st["iterations"] == tr["iterations"] and st["improvements"] == tr["improvements"]
    assert st["degenerate"] == tr["degenerate"]
    return s, M, inl, st


def test_ransac_unit_cases(gpu, hostlib, oracle):
    # test/test_ransac_unit.cpp: empty, too few, identity
    for kind in (0, 1, 2):
        s, M, inl, st = hostlib.ransac(kind, np.zeros((0, 7)))
        assert s == 0 and len(inl) == 0
    sq = np.zeros((4, 7))
    sq[:, 0:3] = [(1, 2, 1), (2, 2, 1), (2, 1, 1), (1, 1, 1)]
    sq[:, 3:6] = sq[:, 0:3]
    s, M, inl, st = check_ransac_equal(hostlib, oracle, 0, sq)
    assert s == pytest.approx(1.0, abs=1e-15) and inl.sum() == 4 and np.linalg.norm(m33(M) - np.eye(3)) < 1e-14
    s, M, inl, st = hostlib.ransac(0, sq[:3])
    assert s == 0 and len(inl) == 3 and not inl.any() and np.isnan(M[:9]).all()
    pts = [(1, 2, 1), (2, 2, 1), (2, 1, 1), (1, 1, 1), (1, 2, 3), (2, 2, 2), (2, 1, 3), (1, 1, 2)]
    c = np.zeros((8, 7))
    c[:, 0:3] = pts
    c[:, 0:3] /= np.linalg.norm(c[:, 0:3], axis=1, keepdims=True)
    c[:, 3:6] = c[:, 0:3]
    s, M, inl, st = check_ransac_equal(hostlib, oracle, 2, c)
    assert s == pytest.approx(1.0, abs=1e-15) and inl.sum() == 8 and abs(np.linalg.norm(m33(M)) - 1) < 1e-14
    s, M, inl, st = check_ransac_equal(hostlib, oracle, 1, c[:6])
    assert s >= 0.16 and inl.sum() >= 1


@pytest.mark.parametrize("name", ["h30", "h80", "hdeg", "f30", "fplane", "e0", "h30q", "c1"])
def test_ransac_equals_reference_driver_golden(gpu, hostlib, golden, name):
    # golden = the reference's own ransac.cpp (driver) over the restated model functions
    kind = int(golden[f"rs_{name}_kind"]) if f"rs_{name}_kind" in golden.files else 0
    s, M, inl, st = hostlib.ransac(kind, golden[f"rs_{name}_corr"])
    assert s == float(golden[f"rs_{name}_score"])
    assert np.array_equal(M[:9], golden[f"rs_{name}_M"][:9])
    assert np.array_equal(inl, golden[f"rs_{name}_inl"])
    assert st["scored"] >= st["iterations"] - st["degenerate"] and st["gpu_calls"] >= 1


def test_ransac_benchmark_floors(gpu, hostlib, oracle):
    # test/test_ransac_benchmark.cpp floors, through the GPU driver
    def pr(inl, n_true):
        gt = np.arange(len(inl)) < n_true
        tp, fp, fn = (inl & gt).sum(), (inl & ~gt).sum(), (~inl & gt).sum()
        return tp / max(tp + fp, 1), tp / max(tp + fn, 1)

    def merr(M, gt):
        a, b = M / np.linalg.norm(M), gt / np.linalg.norm(gt)
        return min(np.linalg.norm(a - b), np.linalg.norm(a + b))

    for n_in, n_out, pmin, rmin in ((200, 0, .99, .99), (140, 60, .90, .85), (80, 120, .80, .70), (40, 160, .70, .60)):
        corr, gt = oracle.scene_homography(n_in, n_out, 42)
        s, M, inl, st = check_ransac_equal(hostlib, oracle, 0, corr)
        p, r = pr(inl, n_in)
        assert p >= pmin and r >= rmin
        if n_out == 0:
            assert merr(m33(M), gt) < 1e-6
    for n_in, n_out, planar, pmin, rmin in ((200, 0, 0.0, .95, .80), (140, 60, 0.0, .85, .70), (200, 0, 0.8, .95, .95)):
        corr, gt = oracle.scene_fundamental(n_in, n_out, planar, 42)
        s, M, inl, st = check_ransac_equal(hostlib, oracle, 2, corr)
        p, r = pr(inl, n_in)
        assert p >= pmin and r >= rmin


def test_ransac_prosac_and_large(gpu, hostlib, oracle):
    corr, _ = oracle.scene_homography(1500, 1000, 3)
    corr[:, 6] = np.random.default_rng(43).uniform(0.05, 0.4, len(corr))  # PROSAC branch
    check_ransac_equal(hostlib, oracle, 0, corr)
    corr, _ = oracle.scene_homography(600, 5400, 4)  # 10 % inliers: thousands of iterations, SPRT rejections
    s, M, inl, st = check_ransac_equal(hostlib, oracle, 0, corr)
    assert st["iterations"] > 1000
    corr, _ = oracle.scene_fundamental(500, 300, 0.5, 9)  # DEGENSAC path
    check_ransac_equal(hostlib, oracle, 2, corr)
    check_ransac_equal(hostlib, oracle, 1, corr)


def test_degensac_through_the_mirror(gpu, hostlib, oracle):
    corr, _ = oracle.scene_fundamental(200, 0, 0.8, 42)
    eo, samples = oracle.hypothesis_stream(2, corr, 4)
    M = oracle.fit(2, corr, samples[0])
    s, inl = oracle.evaluate(2, M, corr)
    a = hostlib.check_degeneracy_f(M, corr, inl)
    b = oracle.check_degeneracy_f(M, corr, inl)
    assert np.array_equal(a[0][:9], b[0][:9]) and np.array_equal(a[1], b[1])


def test_ransac_batch_equals_single_runs(gpu, hostlib, oracle):
    # lock-step batch (one request-table launch per round) against ransac() run job by job: identical models,
    # inlier sets, scores and iteration counts, for all three models, ragged sizes, trivial and PROSAC jobs included
    scenes_h = [oracle.scene_homography(140, 60, 42)[0], oracle.scene_homography(40, 160, 7)[0],
                oracle.scene_homography(700, 300, 11)[0], np.zeros((0, 7)), oracle.scene_homography(3, 0, 1)[0],
                oracle.scene_homography_near_degenerate()[0]]
    q = oracle.scene_homography(300, 100, 5)[0].copy()
    q[:, 6] = np.random.default_rng(43).uniform(0.05, 0.4, len(q))  # PROSAC branch
    scenes_h.append(q)
    scenes_f = [oracle.scene_fundamental(140, 60, 0.0, 42)[0], oracle.scene_fundamental(200, 0, 0.8, 3)[0],
                oracle.scene_fundamental(60, 20, 0.0, 9)[0]]
    for kind, scenes in ((0, scenes_h), (2, scenes_f), (1, scenes_f[:2])):
        batch = hostlib.ransac_batch(kind, scenes, threads=4)
        n = 18 if kind == 0 else 9
        for c, (s, M, inl, st) in zip(scenes, batch):
            s1, M1, inl1, st1 = hostlib.ransac(kind, c)
            assert s == s1 and np.array_equal(inl, inl1[:len(c)])
            assert np.array_equal(M[:n], M1[:n], equal_nan=True)
            assert st["iterations"] == st1["iterations"] and st["improvements"] == st1["improvements"]


def test_device_refit_equals_host_refit(gpu, hostlib, oracle):
    """The lock-step driver refits on the device (OCB_REQ_REFIT_EVALUATE: homography_model::fitInliers,
    homography_model.cpp:52-87, a (2m+1) x 9 full-pivot LU per image pair), the single-run driver on the host; models,
    inlier sets, scores and iteration counts must be identical for every system size, including systems shorter than
    nine rows, all-inlier sets, repeated points and outlier-dominated sets."""
    rng = np.random.default_rng(21)
    scenes = []
    for n_in, n_out, seed in ((4, 0, 1), (5, 1, 2), (9, 0, 3), (33, 31, 4), (257, 3, 5), (2100, 0, 6), (1500, 900, 7),
                              (64, 600, 8)):
        scenes.append(oracle.scene_homography(n_in, n_out, seed)[0])
    dup = oracle.scene_homography(120, 40, 9)[0].copy()
    dup[10:60] = dup[10]                      # fifty copies of one correspondence
    scenes.append(dup)
    noisy = oracle.scene_homography(800, 200, 10)[0].copy()
    noisy[:, 0:2] += rng.normal(0, 2e-4, (len(noisy), 2))
    scenes.append(noisy)
    batch = hostlib.ransac_batch(0, scenes, threads=4)
    for c, (s, M, inl, st) in zip(scenes, batch):
        s1, M1, inl1, st1 = hostlib.ransac(0, c)
        assert s == s1 and np.array_equal(inl, inl1[:len(c)])
        assert np.array_equal(M, M1, equal_nan=True)
        assert st["iterations"] == st1["iterations"] and st["improvements"] == st1["improvements"]
        so, Mo, io, tr = oracle.ransac(0, c)
        assert s == so and np.array_equal(M, Mo, equal_nan=True) and np.array_equal(inl, io)


def test_device_fit_bit_exact(gpu, oracle):
    # K3: checkSampleDegeneracy + fit on the device (homography_model.cpp:19-50,120-136) against the oracle's fit:
    # identical models (H and H^-1), identical degeneracy verdicts; repeated points, collinear triples, rank-deficient
    # systems and non-finite coordinates included
    corr, _ = oracle.scene_homography(300, 200, 17)
    eo, samples = oracle.hypothesis_stream(0, corr, 400)
    samples = [list(s) for s in samples]
    samples += [[5, 5, 6, 7], [1, 2, 3, 1]]                       # repeated correspondence
    near, _ = oracle.scene_homography_near_degenerate()
    corr2 = np.concatenate([corr, near])
    base = len(corr)
    rng = np.random.default_rng(3)
    for _ in range(60):
        samples.append(list(base + rng.choice(len(near), 4, replace=False)))
    line = np.zeros((4, 7))                                        # three collinear source points
    line[:, 0:3] = [(0.1, 0.1, 1), (0.2, 0.2, 1), (0.3, 0.3, 1), (0.5, -0.2, 1)]
    line[:, 3:6] = [(0.1, 0.2, 1), (0.25, 0.2, 1), (0.3, 0.35, 1), (0.5, -0.1, 1)]
    bad = corr[:4].copy()
    bad[0, 2] = 0.0                                                # z == 0 -> inf / nan coordinates
    bad[1, 3] = np.nan
    corr2 = np.concatenate([corr2, line, bad])
    o = len(corr2) - 8
    samples += [[o, o + 1, o + 2, o + 3], [o + 3, o + 2, o + 1, o], [o + 4, o + 5, o + 6, o + 7], [o + 4, 1, 2, 3],
                [o + 5, 1, 2, 3]]
    samples = np.array(samples, np.uint32)
    models, deg = gpu.fit_homography(corr2, samples)
    n_deg = 0
    for s, m, d in zip(samples, models, deg):
        want_deg = bool(oracle.check_sample_degeneracy_h(corr2, s.astype(np.uintp)))
        assert d == want_deg
        if want_deg:
            n_deg += 1
            assert np.isnan(m).all()
        else:
            assert np.array_equal(m, oracle.fit(0, corr2, s.astype(np.uintp)), equal_nan=True)
    assert n_deg >= 2 and n_deg < len(samples) // 2


def test_ransac_device_fit_equals_host_fit(gpu, hostlib, oracle):
    # the same run with the batch fitted on the device (default) and on the host: identical in every output
    scenes = [oracle.scene_homography(140, 60, 42)[0], oracle.scene_homography(40, 160, 7)[0],
              oracle.scene_homography_near_degenerate()[0], oracle.scene_homography(600, 5400, 4)[0]]
    try:
        for corr in scenes:
            hostlib.set_ransac_device_fit(True)
            a = hostlib.ransac(0, corr)
            hostlib.set_ransac_device_fit(False)
            b = hostlib.ransac(0, corr)
            assert a[0] == b[0] and np.array_equal(a[1], b[1], equal_nan=True) and np.array_equal(a[2], b[2])
            assert a[3]["iterations"] == b[3]["iterations"] and a[3]["degenerate"] == b[3]["degenerate"]
        hostlib.set_ransac_device_fit(False)
        bh = hostlib.ransac_batch(0, scenes, threads=2)
        hostlib.set_ransac_device_fit(True)
        bd = hostlib.ransac_batch(0, scenes, threads=2)
        for x, y in zip(bh, bd):
            assert x[0] == y[0] and np.array_equal(x[1], y[1], equal_nan=True) and np.array_equal(x[2], y[2])
    finally:
        hostlib.set_ransac_device_fit(True)
