"""K5 (ratio test + compaction on the device) and K6 (distort_keypoints / image_to_3d on the device), through the C ABI.
References: src/match/match_features.cpp:71-101 (emission order, ratio test), src/distort/distort_keypoints.cpp:48-103.
Bar: survivors, their order and their distances bit-exact against the oracle; rays bit-exact against the numpy oracle
(undistorted branch) and against the C++ mirror's host image_to_3d (both branches, the LM undistortion included)."""
import numpy as np
import pytest

import oc_distort as D
from opencalibration_b200 import synthetic

pytestmark = pytest.mark.gpu


def expected_survivors(oracle, q, c):
    """match_features.cpp:74-97 per query, in query order: (query_k, best_k, best_d) of every emitted match."""
    if len(q) == 0:
        return np.zeros((0, 3), np.int64)
    bk, bd, sd = oracle.match_top2(q, c) if len(c) else (np.zeros(len(q), np.uint32), np.full(len(q), 0xFFFF, np.uint16),
                                                         np.full(len(q), 0xFFFF, np.uint16))
    inf = np.float64(np.inf)
    best = np.where(bd == 0xFFFF, inf, bd.astype(np.float64) * (1.0 / 486))
    second = np.where(sd == 0xFFFF, inf, sd.astype(np.float64) * (1.0 / 486))
    keep = best < 0.8 * second  # :94, IEEE double like the reference
    k = np.nonzero(keep)[0]
    return np.stack([k, bk[k], bd[k]], 1).astype(np.int64)


def test_ratio_compaction_equals_the_reference_rule(gpu, oracle):
    images, pos, pairs = synthetic.grid_survey(3, 3, 900, seed=5)
    images[4] = images[4][:333]  # ragged set sizes
    images[7] = images[7][:0]    # an image without features: as a query set nothing, as a candidate set never a match
    images[2] = images[2][:1]    # a single candidate: second = +inf, every query is emitted (A4)
    for i, d in enumerate(images):
        gpu.register_descriptors(5000 + i, d)
    plist = [(5000 + a, 5000 + b) for a, b in pairs] + [(5000, 5002), (5002, 5000), (5007, 5007)]
    all_pairs = pairs + [(0, 2), (2, 0), (7, 7)]
    cap = sum(len(images[a]) for a, _ in all_pairs)
    out, offs = gpu.match_pairs_ratio(plist, cap)
    assert offs[0] == 0 and offs[-1] == len(out)
    survivors = 0
    for p, (a, b) in enumerate(all_pairs):
        got = out[int(offs[p]):int(offs[p + 1])]
        want = expected_survivors(oracle, images[a], images[b])
        assert np.array_equal(np.stack([got["query_k"], got["best_k"], got["best_d"]], 1).astype(np.int64), want), (p, a, b)
        survivors += len(want)
    assert survivors > 1000
    p_single = len(pairs)  # (0, 2): one candidate row -> all 900 queries emitted in order
    assert offs[p_single + 1] - offs[p_single] == len(images[0])
    with pytest.raises(gpu.OcbError, match="out_capacity"):
        gpu.match_pairs_ratio(plist, 10)
    with pytest.raises(gpu.OcbError):
        gpu.match_pairs_ratio([(5000, 99999)], 100)
    empty, eoffs = gpu.match_pairs_ratio([], 0)
    assert len(empty) == 0 and eoffs.tolist() == [0]
    for i in range(len(images)):
        gpu.unregister_descriptors(5000 + i)


def test_ratio_compaction_ties_and_many_pairs(gpu, oracle):
    # 600 small pairs in one submission (the offsets scan runs over several tiles of 256 pairs) with exact ties:
    # duplicated candidate rows make best == second, which the strict comparison never emits
    rng = np.random.default_rng(8)
    sets = []
    for i in range(40):
        a, b = synthetic.config2_pair(int(rng.integers(1, 70)), int(rng.integers(2, 90)), seed=100 + i)
        if i % 3 == 0:
            b[1::2] = b[::2][:len(b[1::2])]  # every candidate row twice
        sets += [a, b]
    for i, s in enumerate(sets):
        gpu.register_descriptors(6000 + i, s)
    plist = [(int(rng.integers(0, len(sets))), int(rng.integers(0, len(sets)))) for _ in range(600)]
    out, offs = gpu.match_pairs_ratio([(6000 + a, 6000 + b) for a, b in plist], sum(len(sets[a]) for a, _ in plist))
    for p, (a, b) in enumerate(plist):
        got = out[int(offs[p]):int(offs[p + 1])]
        want = expected_survivors(oracle, sets[a], sets[b])
        assert np.array_equal(np.stack([got["query_k"], got["best_k"], got["best_d"]], 1).astype(np.int64), want), p
    for i in range(len(sets)):
        gpu.unregister_descriptors(6000 + i)


def test_device_sort_equals_the_reference_order(gpu, hostlib, oracle):
    """K7: the survivors of every pair in the order the reference's std::sort leaves them (match_features.cpp:100-101,
    unstable, ties included) and their PROSAC ordering (ransac.cpp:83-90): against the oracle's match_features_subset
    (the same libstdc++ std::sort on the same emission order) and the host driver's own ordering."""
    rng = np.random.default_rng(3)
    images, pos, pairs = synthetic.grid_survey(3, 3, 2500, seed=9)
    images[3] = images[3][:40]
    images[6] = images[6][:0]
    for i, d in enumerate(images):
        gpu.register_descriptors(9000 + i, d)
    extra = [(0, 0), (3, 1), (6, 2), (2, 6)]
    all_pairs = pairs + extra
    cap = sum(len(images[a]) for a, _ in all_pairs)
    out, offs, qo = gpu.match_pairs_sorted([(9000 + a, 9000 + b) for a, b in all_pairs], cap)
    unsorted, offs2 = gpu.match_pairs_ratio([(9000 + a, 9000 + b) for a, b in all_pairs], cap)
    assert np.array_equal(offs, offs2)
    ties = 0
    for p, (a, b) in enumerate(all_pairs):
        got = out[int(offs[p]):int(offs[p + 1])]
        n1, n2 = len(images[a]), len(images[b])
        w1, w2, wd = oracle.match_features_subset(images[a], images[b], np.arange(n1, dtype=np.uintp),
                                                  np.arange(n2, dtype=np.uintp)) if n1 and n2 else ([], [], [])
        assert np.array_equal(got["query_k"], w1) and np.array_equal(got["best_k"], w2), (p, a, b)
        assert np.array_equal(got["best_d"] * (1.0 / 486), wd)
        ties += len(got) - len(np.unique(got["best_d"]))
        order = qo[int(offs[p]):int(offs[p + 1])]
        assert np.array_equal(order, hostlib.prosac_order(got["best_d"] * (1.0 / 486))), (p, a, b)
    assert ties > 5000  # the tie order is what makes this more than "sorted by distance"
    same, _, none = gpu.match_pairs_sorted([(9000, 9001)], len(images[0]), want_quality_order=False)
    assert none is None and np.array_equal(same, out[int(offs[0]):int(offs[1])]) if all_pairs[0] == (0, 1) else True
    for i in range(len(images)):
        gpu.unregister_descriptors(9000 + i)


def test_device_sort_of_lists_longer_than_shared_memory(gpu, oracle):
    # 13 000 survivors: more words than fit in the kernel's shared memory -> the global scratch path
    n = 13000
    a, _ = synthetic.config2_pair(n, 8, seed=5)
    rng = np.random.default_rng(1)
    b = a.copy()
    flips = rng.integers(0, 60, n)  # candidate i = query i with a few bits flipped: every query survives the ratio test
    for j in range(60):
        rows = np.nonzero(flips > j)[0]
        bit = rng.integers(0, 486, len(rows))
        b[rows, bit // 64] ^= np.uint64(1) << (bit % 64).astype(np.uint64)
    gpu.register_descriptors(9100, a)
    gpu.register_descriptors(9101, b)
    out, offs, qo = gpu.match_pairs_sorted([(9100, 9101)], n)
    idx = np.arange(n, dtype=np.uintp)
    w1, w2, wd = oracle.match_features_subset(a, b, idx, idx)
    assert len(w1) > 12000
    assert np.array_equal(out["query_k"], w1) and np.array_equal(out["best_k"], w2)
    for s in (9100, 9101):
        gpu.unregister_descriptors(s)


def grid(cols, rows):
    return np.array([(i, j) for i in range(0, cols, cols // 20) for j in range(0, rows, rows // 20)], np.float64)


def test_device_rays_without_distortion_equal_the_oracle(gpu, hostlib):
    # test/test_distort.cpp:34-43 camera; the undistorted branch is bit-exact against the numpy restatement
    p = grid(4000, 3000)
    rays = gpu.image_to_3d(p, hostlib.camera8(6000, (2000, 1500)))
    assert np.array_equal(rays, D.image_to_3d_undistorted(p, 6000, (2000, 1500)))
    q = np.random.default_rng(1).uniform(-500, 6000, (50000, 2))
    cam = hostlib.camera8(5000, (2672, 2008))
    assert np.array_equal(gpu.image_to_3d(q, cam), D.image_to_3d_undistorted(q, 5000, (2672, 2008)))
    assert np.array_equal(gpu.image_to_3d(q, cam), hostlib.image_to_3d(q, cam))
    # ProjectionType::UNKNOWN leaves the ray unset (NaN in the mirror)
    assert np.isnan(gpu.image_to_3d(q[:5], cam, planar=False)).all()


def test_device_rays_with_distortion_equal_the_host_mirror_bit_for_bit(gpu, hostlib):
    # the Levenberg-Marquardt undistortion (distort_keypoints.cpp:74-91) iterates data-dependently; every iterate has
    # to be the host's for the results to be equal
    p = np.concatenate([grid(4000, 3000), np.random.default_rng(2).uniform(-200, 4200, (20000, 2))])
    for radial, tangential in (((0.02, -0.07, 0.1), (0, 0)), ((0.02, -0.07, 0.1), (0.08, -0.08)),
                               ((-0.05, 0, 0), (0, 0)), ((0, 0, 0), (0.01, 0.02)), ((0.3, 0.2, -0.4), (0.05, 0.03))):
        cam = hostlib.camera8(6000, (2000, 1500), radial, tangential)
        dev, host = gpu.image_to_3d(p, cam), hostlib.image_to_3d(p, cam)
        assert np.array_equal(dev, host, equal_nan=True), (radial, tangential)
        back = D.image_from_3d(dev[:400], 6000, (2000, 1500), radial, tangential)
        if radial[0] < 0.3:  # test/test_distort.cpp:45-67: "only solve to 1/100 of a pixel error"
            assert np.abs(back - p[:400]).max() < 1e-2


def test_corr_bind_batch_matches_equals_host_distort_keypoints(gpu, hostlib, oracle):
    survey = synthetic.PlanarSurvey(2, 3, 1200, seed=4)
    imgs = [survey.image(i) for i in range(survey.n_images)]
    cams = [hostlib.camera8(3000.0, (2000, 1500)), hostlib.camera8(2900.0, (1990, 1510), (0.02, -0.01, 0.001), (1e-3, -2e-3)),
            survey.camera8()] * 2
    gpu.register_images_batch([(8000 + i, d, xy, cams[i]) for i, (d, xy, s) in enumerate(imgs)])
    pairs = [(0, 1), (1, 0), (2, 5), (4, 3), (3, 3)]
    out, offs = gpu.match_pairs_ratio([(8000 + a, 8000 + b) for a, b in pairs], sum(len(imgs[a][0]) for a, _ in pairs))
    sets, want = [], []
    rng = np.random.default_rng(0)
    for p, (a, b) in enumerate(pairs):
        m = out[int(offs[p]):int(offs[p + 1])]
        m = m[np.argsort(-m["best_d"].astype(np.int64), kind="stable")]  # some fixed order; row i belongs to match i
        order = rng.permutation(len(m)).astype(np.uint32) if p % 2 == 0 else None
        sets.append((8000 + a, 8000 + b, m, order))
        c = np.zeros((len(m), 7))
        c[:, 0:3] = hostlib.image_to_3d(imgs[a][1][m["query_k"]], cams[a])
        c[:, 3:6] = hostlib.image_to_3d(imgs[b][1][m["best_k"]], cams[b])
        c[:, 6] = m["best_d"] * (1.0 / 486)
        want.append(c)
    sets.append((8000, 8001, np.zeros(0, gpu.MATCH_DTYPE), None))  # an empty set keeps its place in the batch
    got = gpu.corr_bind_batch_matches(sets)
    assert sum(len(w) for w in want) > 1500
    for g, w in zip(got, want):
        assert np.array_equal(g, w)
    assert got[-1].shape == (0, 7)
    bad = out[:4].copy()
    bad["best_k"][2] = 10 ** 6
    with pytest.raises(gpu.OcbError, match="out of range"):
        gpu.corr_bind_batch_matches([(8000, 8001, bad, None)])
    gpu.register_descriptors(8100, imgs[0][0])  # a set without keypoints cannot feed the rays
    with pytest.raises(gpu.OcbError, match="keypoints"):
        gpu.corr_bind_batch_matches([(8100, 8001, out[:4], None)])
    for i in list(range(8000, 8006)) + [8100]:
        gpu.unregister_descriptors(i)
