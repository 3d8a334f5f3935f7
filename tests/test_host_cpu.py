"""CPU tests of the product side that need no GPU: the C-ABI library loads and exports what include/ocb.h declares,
fails loudly without a device, and the host-side pieces of the C++ mirror (layout, fits, error, subsample,
degeneracy tests, linear algebra) agree with the oracle."""
import ctypes as C
import os

import numpy as np
import pytest

import oc_oracle as O
from opencalibration_b200 import capi, synthetic


def test_library_exports_every_declared_symbol(built):
    L = capi.lib()
    names = capi.exported_symbols()
    assert len(names) >= 19
    for n in names:
        assert hasattr(L, n), f"libocb.so does not export {n}"
    assert hasattr(L, "ocb_probe_pipes")  # include/ocb_probe.h


def test_no_cpu_fallback_without_device(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    L = capi.lib()
    assert L.ocb_device_count() == 0
    q = np.zeros((4, 8), np.uint64)
    out = np.zeros(4, capi.TOP2_DTYPE)
    rc = L.ocb_match_top2(q.ctypes.data_as(C.c_void_p), 4, q.ctypes.data_as(C.c_void_p), 4,
                          out.ctypes.data_as(C.c_void_p), None)
    assert rc != 0 and b"no CUDA device" in L.ocb_last_error()
    with pytest.raises(capi.OcbError):
        capi.match_top2(q, q)
    with pytest.raises(capi.OcbError):
        capi.score_models(0, np.zeros((1, 18)), np.zeros((4, 7)), 0.005)


def test_product_never_imports_the_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "opencalibration_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cpp", ".hpp", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oc_oracle" not in text and "oracle/" not in text.replace("the oracle/", ""), f


def test_argument_validation(built):
    L = capi.lib()
    assert L.ocb_set_option(b"no_such_option", 1) == capi.lib().ocb_set_option(b"no_such_option", 1) != 0
    assert L.ocb_score_models(7, None, 1, None, 0, 0.0, None, None, None, None) != 0
    assert L.ocb_match_top2_workspace_bytes(10000, 10000, 1) > L.ocb_match_top2_workspace_bytes(10000, 10000, 0) > 0


def test_mirror_layout(hostlib):
    assert hostlib.layout() == dict(sizeof_feature_2d=96, offsetof_descriptor=24, sizeof_feature_match=24,
                                    sizeof_correspondence=56, sizeof_feature_match_denormalized=64)


def test_mirror_subsample(hostlib, oracle, config1, golden):
    xy = np.array([[100, 100], [110, 105], [200, 200], [205, 202], [300, 300]], float)
    st = np.array([0.5, 0.9, 0.3, 0.7, 0.4], np.float32)
    assert hostlib.spatially_subsample_feature_indices(xy, st, 20.0).tolist() == [1, 3, 4]  # test_match.cpp:90-107
    assert len(hostlib.spatially_subsample_feature_indices(np.zeros((0, 2)), np.zeros(0, np.float32), 20.0)) == 0
    a = hostlib.spatially_subsample_feature_indices(config1["a_xy"], config1["a_strength"], 40.0)
    assert np.array_equal(a, golden["c1_idx_a"])
    b = hostlib.spatially_subsample_feature_indices(config1["b_xy"], config1["b_strength"], 40.0, count=5000)
    assert np.array_equal(b, oracle.subsample(config1["b_xy"], config1["b_strength"], 40.0, count=5000))


def test_mirror_linalg(hostlib, oracle):
    rng = np.random.default_rng(3)
    for rows in (9, 15, 41):
        A, b = rng.standard_normal((rows, 9)), rng.standard_normal(rows)
        assert np.array_equal(hostlib.full_piv_lu_solve(A, b), oracle.fullpivlu_solve(A, b))
    M = rng.standard_normal((3, 3))
    assert np.array_equal(hostlib.invert3(M), oracle.inverse3(M))
    A = rng.standard_normal((9, 9))
    for x, y in zip(hostlib.jacobi_svd(A), oracle.jacobi_svd_square(A)):
        assert np.array_equal(x, y)
    T = rng.standard_normal((17, 3))
    for x, y in zip(hostlib.jacobi_svd_tall(T), oracle.jacobi_svd_tall_v(T)):
        assert np.array_equal(x, y)
    assert np.allclose(hostlib.jacobi_svd(A)[1], np.linalg.svd(A)[1], atol=1e-12)


@pytest.mark.parametrize("kind", [O.KIND_H, O.KIND_E, O.KIND_F])
def test_mirror_fit_and_error_equal_oracle(hostlib, oracle, kind):
    corr, _ = oracle.scene_homography(140, 60) if kind == O.KIND_H else oracle.scene_fundamental(140, 60)
    eo, samples = oracle.hypothesis_stream(kind, corr, 40)
    for s in samples:
        a, b = hostlib.fit(kind, corr, s), oracle.fit(kind, corr, s)
        n = 18 if kind == O.KIND_H else 9
        assert np.array_equal(a[:n], b[:n])
        assert [hostlib.error(kind, a, c) for c in corr[:25]] == [oracle.error(kind, b, c) for c in corr[:25]]
    inl = np.arange(len(corr)) < 140
    M0 = hostlib.fit(kind, corr, samples[0])
    n = 18 if kind == O.KIND_H else 9
    assert np.array_equal(hostlib.fit_inliers(kind, M0, corr, inl)[:n], oracle.fit_inliers(kind, M0, corr, inl)[:n])
    # essential / fundamental fitInliers with too few inliers leaves the model unchanged
    if kind != O.KIND_H:
        few = np.arange(len(corr)) < 3
        assert np.array_equal(hostlib.fit_inliers(kind, M0, corr, few)[:9], M0[:9])


def test_mirror_error_edge_cases(hostlib, oracle):
    corr, _ = oracle.scene_homography(10, 0)
    M = hostlib.fit(O.KIND_H, corr, [0, 1, 2, 3])
    c = corr[5].copy()
    c[2] = 0.0  # measurement1.z == 0 -> NaN residual, like the reference's m / m.z
    assert np.isnan(hostlib.error(O.KIND_H, M, c)) and np.isnan(oracle.error(O.KIND_H, M, c))
    E = np.zeros(18)  # zero matrix: denominator < 1e-20 -> DBL_MAX (essential_matrix_model.cpp:119-120)
    assert hostlib.error(O.KIND_E, E, corr[0]) == np.finfo(np.float64).max == oracle.error(O.KIND_E, E, corr[0])


def test_mirror_sample_degeneracy(hostlib, oracle):
    pts = [(0, 0, 1), (1, 1, 1), (2, 2, 1), (0, 1, 1), (1, 0, 1)]
    c = np.zeros((5, 7))
    c[:, 0:3] = pts
    c[:, 3:6] = pts
    assert hostlib.check_sample_degeneracy_h(c, [0, 1, 2, 3]) and oracle.check_sample_degeneracy_h(c, [0, 1, 2, 3])
    assert not hostlib.check_sample_degeneracy_h(c, [0, 1, 3, 4])
    assert not oracle.check_sample_degeneracy_h(c, [0, 1, 3, 4])


def test_mirror_assemble_inliers(hostlib):
    # ransac.cpp:263-282
    xy1 = np.arange(20, dtype=float).reshape(10, 2)
    xy2 = xy1 + 100
    i1, i2, d = np.array([3, 1, 7, 2]), np.array([0, 9, 4, 4]), np.array([0.4, 0.3, 0.2, 0.1])
    px, ix = hostlib.assemble_inliers(i1, i2, d, [True, False, True, True], xy1, xy2)
    assert ix.tolist() == [[3, 0, 0], [7, 4, 2], [2, 4, 3]]
    assert np.array_equal(px[:, :2], xy1[[3, 7, 2]]) and np.array_equal(px[:, 2:], xy2[[0, 4, 4]])


def test_mirror_essential_decompose(hostlib):
    # essential_matrix_model.cpp:125-153: E = [t]_x R -> one of the four poses reproduces (R, +-t)
    ang = 0.2
    R = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]])
    t = np.array([0.6, 0.0, 0.8])
    tx = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]])
    E = tx @ R
    M18 = np.zeros(18)
    M18[:9] = E.T.ravel()
    poses = hostlib.decompose_essential(M18)
    best = 1e9
    for q in poses:
        x, y, z, w = q[:4]
        Rq = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                       [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                       [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
        assert abs(np.linalg.det(Rq) - 1) < 1e-9
        best = min(best, np.linalg.norm(Rq - R) + min(np.linalg.norm(q[4:] - t), np.linalg.norm(q[4:] + t)))
    assert best < 1e-9


def test_synthetic_generators_are_seeded():
    a1, b1 = synthetic.config2_pair(300, 200, seed=5)
    a2, b2 = synthetic.config2_pair(300, 200, seed=5)
    assert np.array_equal(a1, a2) and np.array_equal(b1, b2)
    assert int(a1[:, 7].max()) < (1 << 38)  # bits 486..511 are zero
    imgs, pos, pairs = synthetic.grid_survey(3, 4, 128, seed=7)
    assert len(imgs) == 12 and all(i.shape == (128, 8) for i in imgs)
    assert all(a != b for a, b in pairs) and len(pairs) == 12 * 9


def test_subsample_equals_the_reference_code_on_adversarial_layouts(hostlib, reference):
    """spatially_subsample_feature_indices (match_features.cpp:8-52) answers the nearest-kept-neighbour query from a hash
    grid and sorts compact records; the decisions and the order must be the reference's own (KD-tree, indirect sort)
    also for points on cell borders, duplicates, strength ties, negative and huge coordinates."""
    rng = np.random.default_rng(0)
    for trial in range(150):
        n = int(rng.integers(1, 400))
        sp = float(rng.choice([0.5, 1.0, 7.3, 40.0, 1e-3, 1e4]))
        xy = rng.integers(-50, 50, (n, 2)) * sp * rng.choice([1.0, 0.5, 1.001, 0.999]) + rng.choice([0, 1e7, -3e8])
        if trial % 3 == 0:
            xy = rng.normal(0, sp * 3, (n, 2))
        st = rng.integers(0, 8, n).astype(np.float32) / 8 if trial % 2 else rng.random(n).astype(np.float32)
        count = 0 if trial % 5 else int(rng.integers(1, n + 1))
        got = hostlib.spatially_subsample_feature_indices(xy, st, sp, count)
        assert np.array_equal(got, reference.subsample(xy, st, sp, count)), (trial, n, sp, count)
