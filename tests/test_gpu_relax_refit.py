"""GPU parity test of the batched refit runner for the second caller of homography fitInliers / evaluate,
RelaxGroup::finalize (reference src/relax/relax_group.cpp:137-177): per edge
    for (int i = 0; i < 3; i++) { h.fitInliers(correspondences, inliers); h.evaluate(correspondences, inliers); }
Bar: model, inliers and score of every edge equal the per-edge loop of the mirror (host fit + device evaluate) and of the
CPU oracle, bit for bit."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

H = 0  # OCB_MODEL_HOMOGRAPHY


def edges(oracle, n_edges):
    """Synthetic edges of a measurement graph: homography scenes of different sizes with the inliers a previous RANSAC
    left (here: the oracle's), some of them perturbed, one with very few inliers, one empty."""
    rng = np.random.default_rng(7)
    out = []
    for e in range(n_edges):
        n_in, n_out = 60 + 37 * e, 20 + 11 * e
        corr, _ = oracle.scene_homography(n_in, n_out, 100 + e)
        _, _, inl, _ = oracle.ransac(H, corr)
        inl = inl.copy()
        if e % 3 == 1:  # the camera model changed: a tenth of the old inliers is lost (evaluate brings them back)
            drop = rng.choice(np.flatnonzero(inl), int(inl.sum()) // 10, replace=False)
            inl[drop] = False
        if e == 7:  # ... and here outliers are flagged as inliers: the least-squares refit is ruined, evaluate finds
            flip = rng.choice(len(inl), len(inl) // 10, replace=False)  # nothing and the next refit sees no inlier
            inl[flip] = ~inl[flip]
        if e % 5 == 4:  # only a handful of old inliers
            keep = np.flatnonzero(inl)[:6]
            inl[:] = False
            inl[keep] = True
        out.append((corr, inl))
    out.append((np.zeros((0, 7)), np.zeros(0, bool)))  # an edge without matches
    return out


def per_edge_loop(fit_inliers, evaluate, corr, inl, rounds=3):
    M = np.full(18, np.nan)
    score = 0.0
    inl = inl.copy()
    for _ in range(rounds):
        M = fit_inliers(H, M, corr, inl)
        score, inl = evaluate(H, M, corr)
    return score, M, inl


def test_batched_refit_equals_the_per_edge_loop(gpu, hostlib, oracle):
    E = edges(oracle, 9)
    got = hostlib.refit_evaluate_batch([c for c, _ in E], [i for _, i in E], rounds=3)
    assert len(got) == len(E)
    good = 0
    for (corr, inl), (score, M, new_inl) in zip(E[:-1], got[:-1]):
        s_m, M_m, i_m = per_edge_loop(hostlib.fit_inliers, hostlib.evaluate, corr, inl)
        assert score == s_m and np.array_equal(M, M_m, equal_nan=True) and np.array_equal(new_inl, i_m)
        s_o, M_o, i_o = per_edge_loop(oracle.fit_inliers, oracle.evaluate, corr, inl)
        assert score == s_o and np.array_equal(M, M_o, equal_nan=True) and np.array_equal(new_inl, i_o)
        good += bool(new_inl.sum() >= 4 and score > 0)
    assert good == len(E) - 2  # every edge but the ruined one (and the empty one) ends with a supported homography
    score, M, new_inl = got[-1]
    assert score == 0 and len(new_inl) == 0


def test_batched_refit_rounds_and_threshold(gpu, hostlib, oracle):
    corr, _ = oracle.scene_homography(300, 120, 5)
    _, _, inl, _ = oracle.ransac(H, corr)
    one = hostlib.refit_evaluate_batch([corr], [inl], rounds=1)[0]
    s, M, i = per_edge_loop(hostlib.fit_inliers, hostlib.evaluate, corr, inl, rounds=1)
    assert one[0] == s and np.array_equal(one[1], M, equal_nan=True) and np.array_equal(one[2], i)
    tight = hostlib.refit_evaluate_batch([corr], [inl], rounds=2, thr=0.001)[0]
    M2 = hostlib.fit_inliers(H, np.full(18, np.nan), corr, inl)
    s2, i2 = hostlib.evaluate(H, M2, corr, thr=0.001)
    M2 = hostlib.fit_inliers(H, M2, corr, i2)
    s2, i2 = hostlib.evaluate(H, M2, corr, thr=0.001)
    assert tight[0] == s2 and np.array_equal(tight[1], M2, equal_nan=True) and np.array_equal(tight[2], i2)
    assert tight[2].sum() <= one[2].sum()
    with pytest.raises(Exception):
        hostlib.refit_evaluate_batch([corr], [inl[:-1]])
