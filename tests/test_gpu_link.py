"""Batched LinkStage runner (host/link_batch.hpp; reference src/pipeline/link_stage.cpp:75-112): every pair's
camera_relations must equal what the per-pair reference flow gives -- subsample, match, rays, RANSAC, decomposition,
inlier assembly -- with the oracle standing in for the reference functions."""
import numpy as np
import pytest

import oc_decompose
import oc_distort
import oc_oracle as O
from opencalibration_b200 import synthetic

pytestmark = pytest.mark.gpu


def expected_pair(oracle, img_a, img_b, cam, spacing=40.0, num_sparse=(0, 0)):
    (da, xa, sa), (db, xb, sb) = img_a, img_b
    ia = oracle.subsample(xa, sa, spacing, num_sparse[0]) if len(da) else np.zeros(0, np.uintp)
    ib = oracle.subsample(xb, sb, spacing, num_sparse[1]) if len(db) else np.zeros(0, np.uintp)
    m1, m2, md = oracle.match_features_subset(da, db, ia, ib)
    corr = np.zeros((len(m1), 7))
    corr[:, 0:3] = oc_distort.image_to_3d_undistorted(xa[m1], cam[0], cam[1:3]) if len(m1) else 0
    corr[:, 3:6] = oc_distort.image_to_3d_undistorted(xb[m2], cam[0], cam[1:3]) if len(m1) else 0
    corr[:, 6] = md
    score, M, inl, _ = oracle.ransac(O.KIND_H, corr)
    H = M[:9].reshape(3, 3).T
    ok, poses = oc_decompose.homography_decompose(H, corr, inl) if not np.isnan(H).any() else (False, None)
    keep = ok and inl.sum() > 4 * 1.5
    return dict(matches=(m1, m2, md), H=H, inl=inl, ok=ok, keep=keep, poses=poses, xa=xa, xb=xb)


def check_pair(got, exp):
    assert np.array_equal(got["H"], exp["H"], equal_nan=True)
    assert got["relation_type"] == 0  # RelationType::HOMOGRAPHY (link_stage.cpp:96)
    if exp["poses"] is not None:
        assert np.array_equal(got["poses"][:, 7], exp["poses"][:, 7])
        assert np.allclose(got["poses"][:, :7], exp["poses"][:, :7], atol=1e-9, rtol=0, equal_nan=True)
    m1, m2, md = exp["matches"]
    if exp["keep"]:
        assert np.array_equal(got["matches"][0], m1) and np.array_equal(got["matches"][1], m2)
        assert np.array_equal(got["matches"][2], md)
        idx = np.nonzero(exp["inl"])[0]
        assert np.array_equal(got["inlier_idx"][:, 2], idx)
        assert np.array_equal(got["inlier_idx"][:, 0], m1[idx]) and np.array_equal(got["inlier_idx"][:, 1], m2[idx])
        assert np.array_equal(got["inlier_pixels"][:, 0:2], exp["xa"][m1[idx]])
        assert np.array_equal(got["inlier_pixels"][:, 2:4], exp["xb"][m2[idx]])
    else:  # link_stage.cpp:104: matches are only stored for decomposable, well-supported relations
        assert len(got["matches"][0]) == 0 and len(got["inlier_idx"]) == 0


def test_config1_pair_both_directions(gpu, hostlib, oracle, config1):
    imgs = [(config1["a_desc"], config1["a_xy"], config1["a_strength"]),
            (config1["b_desc"], config1["b_xy"], config1["b_strength"])]
    cam = hostlib.camera8(5000, (2672, 2008))  # test/test_ransac_functional.cpp:26-31
    sets = [hostlib.FeatureSet(d, xy, s) for d, xy, s in imgs]
    res = hostlib.link_pairs(sets, [cam, cam], [(0, 1), (1, 0)], threads=2)
    for p, (a, b) in enumerate([(0, 1), (1, 0)]):
        exp = expected_pair(oracle, imgs[a], imgs[b], cam)
        got = res.get(p)
        check_pair(got, exp)
        assert exp["keep"] and len(got["matches"][0]) > 400 and len(got["inlier_idx"]) > 100
    assert res.stats["comparisons"] == 2 * 4052 * 4198


def test_planar_survey_grid(gpu, hostlib, oracle):
    survey = synthetic.PlanarSurvey(3, 3, 1500, seed=11)
    imgs = [survey.image(i) for i in range(survey.n_images)]
    imgs[4] = tuple(x[:400] for x in imgs[4])  # ragged
    imgs[8] = tuple(x[:0] for x in imgs[8])    # an image without features
    cam = survey.camera8()
    sets = [hostlib.FeatureSet(d, xy, s) for d, xy, s in imgs]
    pairs = survey.pairs + [(2, 2)]
    num_sparse = [0] * 9
    num_sparse[1] = 700  # only the sparse prefix of image 1 is subsampled (link_stage.cpp:63-65)
    res = hostlib.link_pairs(sets, [cam] * 9, pairs, num_sparse=num_sparse, threads=4, pairs_per_submission=7)
    kept = 0
    for p, (a, b) in enumerate(pairs):
        exp = expected_pair(oracle, imgs[a], imgs[b], cam, num_sparse=(num_sparse[a], num_sparse[b]))
        check_pair(res.get(p), exp)
        kept += bool(exp["keep"])
    assert kept >= 20  # neighbouring images genuinely overlap
    # matches only
    res2 = hostlib.link_pairs(sets, [cam] * 9, pairs[:5], num_sparse=num_sparse, run_ransac=False)
    for p, (a, b) in enumerate(pairs[:5]):
        exp = expected_pair(oracle, imgs[a], imgs[b], cam, num_sparse=(num_sparse[a], num_sparse[b]))
        got = res2.get(p)
        assert np.array_equal(got["matches"][0], exp["matches"][0]) and np.array_equal(got["matches"][2], exp["matches"][2])


def test_link_pairs_rejects_bad_input(gpu, hostlib):
    survey = synthetic.PlanarSurvey(1, 2, 100, seed=3)
    sets = [hostlib.FeatureSet(*survey.image(i)) for i in range(2)]
    with pytest.raises(hostlib.OcbError):
        hostlib.link_pairs(sets, [survey.camera8()] * 2, [(0, 5)])
    # a sparse-feature count beyond the feature vector (a corrupt checkpoint) is rejected, not read out of bounds
    with pytest.raises(hostlib.OcbError):
        hostlib.link_pairs(sets, [survey.camera8()] * 2, [(0, 1)], num_sparse=[101, 0])


def test_link_pairs_edge_cases(gpu, hostlib, oracle):
    """Submission sizes around the list length, an empty list, repeated and self pairs, images without features on
    either side, a failing call between two good ones, and a submission size of one with two submissions in flight:
    every result equals the per-pair flow and the runner stays usable."""
    survey = synthetic.PlanarSurvey(2, 3, 600, seed=21)
    imgs = [survey.image(i) for i in range(survey.n_images)]
    imgs[5] = tuple(x[:0] for x in imgs[5])  # no features at all
    cam = survey.camera8()
    cams = [cam] * survey.n_images
    sets = [hostlib.FeatureSet(d, xy, s) for d, xy, s in imgs]
    pairs = survey.pairs + [(1, 1), (0, 1), (0, 1), (5, 0), (0, 5), (5, 5)]
    want = [expected_pair(oracle, imgs[a], imgs[b], cam) for a, b in pairs]

    def check(res, which):
        assert res.n_pairs == len(which)
        for p, k in enumerate(which):
            check_pair(res.get(p), want[k])

    everything = list(range(len(pairs)))
    # one submission far larger than the list, exactly the list, one pair per submission (two in flight), and a size
    # that leaves a ragged last submission
    for pps in (10_000, len(pairs), 1, len(pairs) - 1):
        check(hostlib.link_pairs(sets, cams, pairs, threads=3, pairs_per_submission=pps), everything)
    # an empty pair list is a valid call
    empty = hostlib.link_pairs(sets, cams, [], threads=2)
    assert empty.n_pairs == 0 and empty.stats["comparisons"] == 0
    # a single pair, a single worker thread
    check(hostlib.link_pairs(sets, cams, pairs[:1], threads=1, pairs_per_submission=1), [0])
    # a rejected call (index out of range in the middle of the list) leaves nothing behind: the next call is complete
    bad = list(pairs)
    bad[len(bad) // 2] = (0, 99)
    with pytest.raises(hostlib.OcbError):
        hostlib.link_pairs(sets, cams, bad, threads=3, pairs_per_submission=2)
    check(hostlib.link_pairs(sets, cams, pairs, threads=3, pairs_per_submission=2), everything)
    # matches only (no RANSAC): images without features give empty lists, repeated pairs equal lists
    res = hostlib.link_pairs(sets, cams, pairs, run_ransac=False, pairs_per_submission=3)
    for p in range(len(pairs)):
        got, exp = res.get(p), want[p]
        assert np.array_equal(got["matches"][0], exp["matches"][0]) and np.array_equal(got["matches"][1], exp["matches"][1])
        assert np.array_equal(got["matches"][2], exp["matches"][2])
    assert len(res.get(len(pairs) - 1)["matches"][0]) == 0 and len(res.get(len(pairs) - 3)["matches"][0]) == 0


def test_c4_survey_slice_equals_the_reference_object_code(gpu, hostlib):
    """BASELINE configs[3] at full image size: >= 200 directed pairs of the 25 x 40 survey, 8192 features per image,
    through the batched runner; EVERY pair's match list (indices, distances, order) against the reference's own
    compiled match_features.cpp (oracle/_ref, -mpopcnt build of the same translation unit), with the subsample
    indices from the reference's own spatially_subsample_feature_indices."""
    import os
    from concurrent.futures import ThreadPoolExecutor
    if not O.Reference.available():
        pytest.skip("oracle/_ref not built")
    ref = O.Reference(popcnt=True)
    survey = synthetic.PlanarSurvey(25, 40, 8192, seed=7)
    pairs = survey.pairs[:225]
    used = sorted({i for p in pairs for i in p})
    local = {g: k for k, g in enumerate(used)}
    with ThreadPoolExecutor(8) as ex:
        imgs = list(ex.map(survey.image, used))
    sets = [hostlib.FeatureSet(d, xy, s) for d, xy, s in imgs]
    lp = [(local[a], local[b]) for a, b in pairs]
    res = hostlib.link_pairs(sets, [survey.camera8()] * len(sets), lp, run_ransac=False, spacing=0.5,
                             pairs_per_submission=64)
    assert res.stats["comparisons"] > 200 * 8000 * 8000
    idx = [ref.subsample(xy, s, 0.5) for _, xy, s in imgs]
    assert all(len(i) > 8000 for i in idx)

    def want(p):
        a, b = lp[p]
        return ref.match_features_subset(imgs[a][0], imgs[b][0], idx[a], idx[b])

    with ThreadPoolExecutor(os.cpu_count() or 4) as ex:  # ctypes releases the GIL: one reference call per core
        expected = list(ex.map(want, range(len(lp))))
    total = 0
    for p, w in enumerate(expected):
        got = res.get(p)["matches"]
        assert all(np.array_equal(x, y) for x, y in zip(got, w)), p
        total += len(w[0])
    assert total > 225 * 500  # neighbouring images share most of their footprint
    res.close()


def _relations_equal(a, b):
    for k in ("H", "poses"):
        if not np.array_equal(a[k], b[k], equal_nan=True):
            return False
    return (a["relation_type"] == b["relation_type"] and all(np.array_equal(x, y) for x, y in zip(a["matches"], b["matches"]))
            and np.array_equal(a["inlier_idx"], b["inlier_idx"]) and np.array_equal(a["inlier_pixels"], b["inlier_pixels"]))


@pytest.mark.parametrize("distorted", [False, True])
def test_device_tail_equals_host_tail(gpu, hostlib, distorted):
    """Ratio test + compaction (K5) and rays (K6) on the device give the relations of the host tail bit for bit -- also
    with a distorted camera model, where every Levenberg-Marquardt iterate of the undistortion has to agree."""
    survey = synthetic.PlanarSurvey(3, 4, 2500, seed=21)
    imgs = [survey.image(i) for i in range(survey.n_images)]
    imgs[5] = tuple(x[:3] for x in imgs[5])   # fewer matches than MINIMUM_POINTS: the run is over before it starts
    imgs[7] = tuple(x[:0] for x in imgs[7])
    cam = survey.camera8()
    if distorted:
        cam = hostlib.camera8(cam[0], cam[1:3], (0.02, -0.03, 0.01), (1e-3, -5e-4))
    sets = [hostlib.FeatureSet(d, xy, s) for d, xy, s in imgs]
    pairs = survey.pairs
    kw = dict(threads=4, pairs_per_submission=16)
    dev = hostlib.link_pairs(sets, [cam] * len(sets), pairs, device_tail=True, **kw)
    ref = hostlib.link_pairs(sets, [cam] * len(sets), pairs, device_tail=False, **kw)
    kept = 0
    for p in range(len(pairs)):
        a, b = dev.get(p), ref.get(p)
        assert _relations_equal(a, b), (p, pairs[p])
        kept += len(a["matches"][0]) > 0
    assert kept >= 40
    assert dev.stats["matches"] == ref.stats["matches"] and dev.stats["ransac_inliers"] == ref.stats["ransac_inliers"]
    # ... and with the two std::sort calls replayed on the device as well (K7: match order and PROSAC order)
    import os
    os.environ["OCB_LINK_DEVICE_SORT"] = "1"
    try:
        srt = hostlib.link_pairs(sets, [cam] * len(sets), pairs, device_tail=True, **kw)
    finally:
        del os.environ["OCB_LINK_DEVICE_SORT"]
    for p in range(len(pairs)):
        assert _relations_equal(srt.get(p), ref.get(p)), (p, pairs[p])
    # matches only
    dev2 = hostlib.link_pairs(sets, [cam] * len(sets), pairs, device_tail=True, run_ransac=False, **kw)
    ref2 = hostlib.link_pairs(sets, [cam] * len(sets), pairs, device_tail=False, run_ransac=False, **kw)
    for p in range(len(pairs)):
        assert all(np.array_equal(x, y) for x, y in zip(dev2.get(p)["matches"], ref2.get(p)["matches"]))
    # the flat copy the tail workers write while the call runs (LinkOptions::packed_out): per pair an offset and a count
    rec = np.zeros(3 * dev2.stats["matches"], np.uint32)
    off, cnt = np.zeros(len(pairs), np.uint64), np.zeros(len(pairs), np.uint64)
    dev3 = hostlib.link_pairs(sets, [cam] * len(sets), pairs, run_ransac=False, packed=(rec, off, cnt), **kw)
    spans = sorted((int(o), int(c)) for o, c in zip(off, cnt) if c)
    assert all(a + n <= b for (a, n), (b, _) in zip(spans, spans[1:])) and int(cnt.sum()) == dev3.stats["matches"]
    for p in range(len(pairs)):
        i1, i2, d = dev3.get(p)["matches"]
        r = rec.reshape(-1, 3)[int(off[p]):int(off[p]) + int(cnt[p])]
        assert np.array_equal(r[:, 0], i1) and np.array_equal(r[:, 1], i2) and np.array_equal(r[:, 2] * (1.0 / 486), d)
    with pytest.raises(hostlib.OcbError, match="packed_capacity"):
        hostlib.link_pairs(sets, [cam] * len(sets), pairs, run_ransac=False, packed=(rec[:30], off, cnt), **kw)
    # with RANSAC the packed lists are the FINAL ones: only pairs that keep a relation keep their matches (:103-107)
    rec[:] = 0
    dev4 = hostlib.link_pairs(sets, [cam] * len(sets), pairs, packed=(rec, off, cnt), **kw)
    for p in range(len(pairs)):
        i1, i2, d = dev4.get(p)["matches"]
        r = rec.reshape(-1, 3)[int(off[p]):int(off[p]) + int(cnt[p])]
        assert len(r) == len(i1) and np.array_equal(r[:, 0], i1) and np.array_equal(r[:, 1], i2)
    # the flat form the ranks gather: counts and 12-byte records, pair after pair
    counts, rec = dev2.pack_matches()
    assert counts.sum() == len(rec) == dev2.stats["matches"]
    o = 0
    for p in range(len(pairs)):
        i1, i2, d = dev2.get(p)["matches"]
        r = rec[o:o + int(counts[p])]
        assert np.array_equal(r[:, 0], i1) and np.array_equal(r[:, 1], i2) and np.array_equal(r[:, 2] * (1.0 / 486), d)
        o += int(counts[p])


def test_link_pairs_over_two_devices_equals_one(gpu, hostlib):
    """In-process multi-GPU runner (host/partition.hpp: link_pairs_multi): the pair list partitioned over two GPUs of this
    process gives the relations of the single-device run, in pair order."""
    if gpu.lib().ocb_device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    survey = synthetic.PlanarSurvey(4, 6, 2000, seed=9)
    imgs = [survey.image(i) for i in range(survey.n_images)]
    sets = [hostlib.FeatureSet(d, xy, s) for d, xy, s in imgs]
    cams = [survey.camera8()] * len(sets)
    one = hostlib.link_pairs(sets, cams, survey.pairs, threads=8, pairs_per_submission=32)
    two = hostlib.link_pairs(sets, cams, survey.pairs, threads=8, pairs_per_submission=32, n_devices=2,
                             positions=survey.positions)
    for p in range(len(survey.pairs)):
        assert _relations_equal(one.get(p), two.get(p)), p
    assert one.stats["matches"] == two.stats["matches"] and one.stats["comparisons"] == two.stats["comparisons"]
    with pytest.raises(hostlib.OcbError):
        hostlib.link_pairs(sets, cams, survey.pairs, n_devices=64, positions=survey.positions)
