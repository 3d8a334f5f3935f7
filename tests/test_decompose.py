"""homography_model::decompose (reference src/model_inliers/homography_model.cpp:138-185): the oracle's restatement
of cv::decomposeHomographyMat against golden vectors from the real OpenCV, and the C++ mirror against both. Host-only
code: no GPU needed."""
import os

import numpy as np
import pytest

import oc_decompose as D
import oc_oracle as O

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "decompose_h.npz"))
TOL = 1e-9  # stated tolerance: the normalisation uses a different SVD than OpenCV's


def test_oracle_decomposition_equals_opencv_golden():
    for H, k, R, t, n in zip(GOLD["H"], GOLD["k"], GOLD["R"], GOLD["t"], GOLD["n"]):
        sol = D.decompose_homography_mat(H)
        assert len(sol) == k
        for j, (Rj, tj, nj) in enumerate(sol):
            assert np.abs(Rj - R[j]).max() < TOL and np.abs(tj - t[j]).max() < TOL and np.abs(nj - n[j]).max() < TOL


def test_oracle_stable_sort_equals_libstdcxx_golden():
    for scores, perm in zip(GOLD["sort_scores"], GOLD["sort_perm"]):
        assert D.stable_sort4(list(scores)) == list(perm), scores


def test_mirror_decomposition_equals_opencv_golden(hostlib):
    for H, k, R, t, n in zip(GOLD["H"], GOLD["k"], GOLD["R"], GOLD["t"], GOLD["n"]):
        Rs, ts, ns = hostlib.decompose_homography_mat(H)
        assert len(Rs) == k
        assert np.abs(Rs - R[:k]).max() < TOL and np.abs(ts - t[:k]).max() < TOL and np.abs(ns - n[:k]).max() < TOL


def _m18(H):
    return np.concatenate([np.asarray(H).T.ravel(), np.linalg.inv(H).T.ravel()])


def test_mirror_decompose_equals_oracle(hostlib, oracle):
    rng = np.random.default_rng(5)
    corr, _ = oracle.scene_homography(140, 60, 42)
    for i in range(0, len(GOLD["H"]), 7):
        H = GOLD["H"][i]
        inl = rng.random(len(corr)) < 0.7
        ok, poses = hostlib.decompose_homography(_m18(H), corr, inl)
        oko, poses_o = D.homography_decompose(H, corr, inl)
        assert ok == oko
        assert np.array_equal(poses[:, 7], poses_o[:, 7])  # scores and their order (ties included)
        assert np.allclose(poses[:, :7], poses_o[:, :7], atol=TOL, rtol=0, equal_nan=True)


def test_reference_unit_cases(hostlib, oracle):
    # test/test_ransac_unit.cpp:22-52 -- identity homography: one solution, zero translation, identity rotation
    corr = np.zeros((4, 7))
    for r, p in enumerate([(1, 2, 1), (2, 2, 1), (2, 1, 1), (1, 1, 1)]):
        corr[r, 0:3] = corr[r, 3:6] = p
    ok, poses = hostlib.decompose_homography(_m18(np.eye(3)), corr, np.ones(4, bool))
    assert ok and poses[0, 7] == 4 and np.all(poses[1:, 7] == -1)
    assert np.linalg.norm(poses[0, 4:7]) < 1e-14
    assert abs(2 * np.arccos(min(1.0, abs(poses[0, 3])))) < 1e-14
    assert np.all(np.isnan(poses[1:, :7]))  # unused slots keep decomposed_pose's NaN defaults

    # test/test_ransac_unit.cpp:114-176 -- plane seen from two poses; one of the four solutions is the true motion
    def quat(axis, ang):
        axis = np.asarray(axis, float) / np.linalg.norm(axis)
        return np.concatenate([axis * np.sin(ang / 2), [np.cos(ang / 2)]])  # x y z w

    def qmul(a, b):
        ax, ay, az, aw = a
        bx, by, bz, bw = b
        return np.array([aw * bx + ax * bw + ay * bz - az * by, aw * by - ax * bz + ay * bw + az * bx,
                         aw * bz + ax * by - ay * bx + az * bw, aw * bw - ax * bx - ay * by - az * bz])

    def qinv(q):
        return np.array([-q[0], -q[1], -q[2], q[3]])

    def qrot(q, v):
        return qmul(qmul(q, np.concatenate([v, [0]])), qinv(q))[:3]

    def angle(q):
        return 2 * np.arctan2(np.linalg.norm(q[:3]), abs(q[3]))

    down = quat([1, 0, 0], np.pi)

    def perspective(v, R, T):
        ray = qrot(qinv(R), v - (np.array([0, 0, 10.0]) + T))
        return np.array([ray[0] / ray[2] * 600, ray[1] / ray[2] * 600, 1.0])

    def normalized(v):  # Eigen's normalized(): a zero vector stays zero
        nrm = np.linalg.norm(v)
        return v / nrm if nrm > 0 else v

    rots = [quat([0, 0, 1], 0.0), quat([0, 0, 1], -np.pi / 2)]  # the reference's parameter grid (:350-359)
    trans = [np.array(t, float) for t in ([0, 0, 0], [1, 0, 0], [1, -1, 0], [-1, 1, 0], [-1, -1, 0])]
    for R in rots:
        for T in trans:
            c = np.zeros((4, 7))
            k = 0
            for i in range(2):
                for j in range(2):
                    p = np.array([-1.0 if i > 0 else 1.0, -1.0 if j > 0 else 1.0, 0.0])
                    c[k, 0:3] = perspective(p, qmul(R, down), T)
                    c[k, 3:6] = perspective(p, down, np.zeros(3))
                    k += 1
            M = oracle.fit(O.KIND_H, c, np.arange(4, dtype=np.uintp))
            M2 = hostlib.fit(O.KIND_H, c, np.arange(4, dtype=np.uintp))
            assert np.allclose(M, M2, rtol=1e-9, atol=1e-12)
            ok, poses = hostlib.decompose_homography(M2, c, np.ones(4, bool))
            assert ok
            errs = []
            for ps in poses:
                if ps[7] < 0:
                    continue
                t_err = np.linalg.norm(qrot(down, normalized(ps[4:7])) - normalized(T))
                r_err = angle(qmul(qinv(qmul(qmul(down, ps[0:4]), qinv(down))), R))
                errs.append(t_err + r_err)
            assert min(errs) < 1e-7, (R, T, errs)
