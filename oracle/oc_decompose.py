"""TEST INFRASTRUCTURE ONLY -- numpy restatement of homography_model::decompose
(reference src/model_inliers/homography_model.cpp:138-185). Only tests/ import this.

The algebra of the reference lives in a third-party call, cv::decomposeHomographyMat (:146; OpenCV is found with
find_package, not vendored: CMakeLists.txt:40, unpinned; CI installs Ubuntu 24.04's libopencv-dev = 4.6). Its
published algorithm (Malis & Vargas, INRIA RR-6303, the analytical decomposition from S = H^T H - I) is restated in
decompose_homography_mat() and PINNED against the outputs of the real cv2.decomposeHomographyMat of this container
(tests/golden/decompose_h.npz, made by tests/golden/make_decompose_vectors.py) to 1e-9 -- not bit-exact, because the
normalisation divides by the middle singular value of OpenCV's own SVD.
What the reference does with the solutions (:150-184) is restated in homography_decompose(): cheirality vote over the
inliers, Eigen::Quaterniond(R), score -1 for unused slots, std::stable_sort with the comparator `score >= score`
(libstdc++'s algorithm for 4 elements restated in stable_sort4 and pinned by the golden table of the same file).
"""
import numpy as np


def _opp_minor(M, row, col):
    x1, x2 = (1 if col == 0 else 0), (1 if col == 2 else 2)
    y1, y2 = (1 if row == 0 else 0), (1 if row == 2 else 2)
    return M[y1, x2] * M[y2, x1] - M[y1, x1] * M[y2, x2]


def _sg(x):
    return 1 if x >= 0 else -1


def decompose_homography_mat(H):
    """-> list of (R, t, n); 1 entry for a pure rotation, else 4 in OpenCV's order."""
    H = np.asarray(H, np.float64)
    W = np.linalg.svd(H, compute_uv=False)
    Hn = H * (1.0 / W[1])
    S = Hn.T @ Hn - np.eye(3)
    if np.abs(S).sum(axis=1).max() < 0.001:
        return [(Hn, np.zeros(3), np.zeros(3))]
    M00, M11, M22 = _opp_minor(S, 0, 0), _opp_minor(S, 1, 1), _opp_minor(S, 2, 2)
    # minors that are zero in exact arithmetic may round to -1e-17: clamp (OpenCV would produce NaN there)
    r00, r11, r22 = np.sqrt(max(M00, 0.0)), np.sqrt(max(M11, 0.0)), np.sqrt(max(M22, 0.0))
    e12, e02, e01 = _sg(_opp_minor(S, 1, 2)), _sg(_opp_minor(S, 0, 2)), _sg(_opp_minor(S, 0, 1))
    a = [abs(S[0, 0]), abs(S[1, 1]), abs(S[2, 2])]
    idx = 0
    if a[0] < a[1]:
        idx = 1
        if a[1] < a[2]:
            idx = 2
    elif a[0] < a[2]:
        idx = 2
    if idx == 0:
        npa = np.array([S[0, 0], S[0, 1] + r22, S[0, 2] + e12 * r11])
        npb = np.array([S[0, 0], S[0, 1] - r22, S[0, 2] - e12 * r11])
    elif idx == 1:
        npa = np.array([S[0, 1] + r22, S[1, 1], S[1, 2] - e02 * r00])
        npb = np.array([S[0, 1] - r22, S[1, 1], S[1, 2] + e02 * r00])
    else:
        npa = np.array([S[0, 2] + e01 * r11, S[1, 2] + r00, S[2, 2]])
        npb = np.array([S[0, 2] - e01 * r11, S[1, 2] - r00, S[2, 2]])
    tr = S[0, 0] + S[1, 1] + S[2, 2]
    v = 2.0 * np.sqrt(max(1 + tr - M00 - M11 - M22, 0.0))
    es = _sg(S[idx, idx])
    r, nt = np.sqrt(2 + tr + v), np.sqrt(max(2 + tr - v, 0.0))
    na, nb = npa / np.linalg.norm(npa), npb / np.linalg.norm(npb)
    ta_s = 0.5 * nt * (es * r * nb - nt * na)
    tb_s = 0.5 * nt * (es * r * na - nt * nb)

    def rot(ts, n):
        R = Hn @ (np.eye(3) - (2 / v) * np.outer(ts, n))
        return -R if np.linalg.det(R) < 0 else R

    Ra, Rb = rot(ta_s, na), rot(tb_s, nb)
    ta, tb = Ra @ ta_s, Rb @ tb_s
    return [(Ra, ta, na), (Ra, -ta, -na), (Rb, tb, nb), (Rb, -tb, -nb)]


def quaternion_from_rotation(R):
    """Eigen::Quaterniond(Matrix3d) -> coeffs (x, y, z, w)."""
    q = np.zeros(4)
    t = R[0, 0] + R[1, 1] + R[2, 2]
    if t > 0:
        t = np.sqrt(t + 1.0)
        q[3] = 0.5 * t
        t = 0.5 / t
        q[0], q[1], q[2] = (R[2, 1] - R[1, 2]) * t, (R[0, 2] - R[2, 0]) * t, (R[1, 0] - R[0, 1]) * t
    else:
        i = 0
        if R[1, 1] > R[0, 0]:
            i = 1
        if R[2, 2] > R[i, i]:
            i = 2
        j, k = (i + 1) % 3, (i + 2) % 3
        t = np.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0)
        q[i] = 0.5 * t
        t = 0.5 / t
        q[3] = (R[k, j] - R[j, k]) * t
        q[j] = (R[j, i] + R[i, j]) * t
        q[k] = (R[k, i] + R[i, k]) * t
    return q


def stable_sort4(scores):
    """Permutation std::stable_sort (libstdc++, g++ 13) produces on 4 elements with comp(a, b) = a.score >= b.score:
    __stable_sort_adaptive sorts the halves [0,2) and [2,4) with an insertion sort, then __move_merge_adaptive merges
    them taking from the SECOND run whenever comp(second, first) (bits/stl_algo.h)."""
    comp = lambda a, b: scores[a] >= scores[b]  # noqa: E731
    runs = []
    for lo in (0, 2):
        a, b = lo, lo + 1
        runs.append([b, a] if comp(b, a) else [a, b])  # insertion sort of two: *i moves to the front if comp(*i, *first)
    first, second, out = runs[0], runs[1], []
    while first and second:
        if comp(second[0], first[0]):
            out.append(second.pop(0))
        else:
            out.append(first.pop(0))
    return out + first + second


def homography_decompose(H, corr, inliers):
    """-> (ok, poses [4][8] = qx,qy,qz,qw, tx,ty,tz, score), poses sorted like the reference sorts them; slots
    beyond the number of solutions keep NaN pose and score -1."""
    corr = np.asarray(corr, np.float64).reshape(-1, 7)
    inl = np.asarray(inliers, bool)
    poses = np.full((4, 8), np.nan)
    poses[:, 7] = -1
    for i, (R, t, n) in enumerate(decompose_homography_mat(H)):
        m1, m2 = corr[inl, 0:3], corr[inl, 3:6]
        score = int(np.sum((m1 @ n >= 0) & (m2 @ (R @ n) >= 0)))
        poses[i, 0:4], poses[i, 4:7], poses[i, 7] = quaternion_from_rotation(R), t, score
    poses = poses[stable_sort4(poses[:, 7])]
    return bool(poses[0, 7] > 0), poses
