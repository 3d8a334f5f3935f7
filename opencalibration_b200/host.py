"""ctypes binding of the C++ mirror's flat shim (libocb_host.so): the reference's entry points
(match_features_subset, spatially_subsample_feature_indices, ransac<Model>, Model::fit/fitInliers/evaluate/error,
assembleInliers) as a Python caller sees them. Everything bulk runs on the GPU through libocb.so; no CPU fallback."""
import ctypes as C
import os

import numpy as np

from .capi import OcbError, PKG, lib as _cuda_lib

KIND_H, KIND_E, KIND_F = 0, 1, 2
MIN_POINTS = {KIND_H: 4, KIND_E: 5, KIND_F: 8}

_u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_szp = np.ctypeslib.ndpointer(np.uintp, flags="C_CONTIGUOUS")

_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    _cuda_lib()  # libocb.so first (also the loud failure when it is missing)
    path = os.path.join(PKG, "libocb_host.so")
    if not os.path.exists(path):
        raise OcbError(f"{path} not found: run `python -m opencalibration_b200.build`")
    L = C.CDLL(path)
    sz, i32, dbl, vp = C.c_size_t, C.c_int, C.c_double, C.c_void_p
    L.ocbh_last_error.restype = C.c_char_p
    for f in ("ocbh_sizeof_feature_2d", "ocbh_offsetof_descriptor", "ocbh_sizeof_feature_match",
              "ocbh_sizeof_correspondence", "ocbh_sizeof_feature_match_denormalized"):
        getattr(L, f).restype = sz
    L.ocbh_match_features_subset.argtypes = [_u64p, sz, _u64p, sz, _szp, sz, _szp, sz, _szp, _szp, _f64p, vp, _szp]
    L.ocbh_features_create.argtypes = [vp, vp, _u64p, sz]
    L.ocbh_features_create.restype = vp
    L.ocbh_features_destroy.argtypes = [vp]
    L.ocbh_features_destroy.restype = None
    L.ocbh_match_handles.argtypes = [vp, vp, _szp, sz, _szp, sz, _szp, _szp, _f64p, vp, _szp]
    L.ocbh_match_guided.argtypes = [_u64p, sz, _u64p, sz, _szp, _szp, _szp, sz, _szp, _szp, _szp, _f64p, _f64p, _szp]
    L.ocbh_subsample.argtypes = [_f64p, _f32p, sz, dbl, sz, _szp]
    L.ocbh_subsample.restype = sz
    L.ocbh_ransac.argtypes = [i32, _f64p, sz, _f64p, _u8p, _f64p, _szp]
    L.ocbh_evaluate.argtypes = [i32, _f64p, dbl, _f64p, sz, _u8p, _f64p]
    L.ocbh_set_ransac_device_fit.argtypes = [i32]
    L.ocbh_set_ransac_device_fit.restype = None
    L.ocbh_error.argtypes = [i32, _f64p, _f64p]
    L.ocbh_error.restype = dbl
    L.ocbh_fit.argtypes = [i32, _f64p, sz, _szp, _f64p]
    L.ocbh_fit.restype = None
    L.ocbh_fit_inliers.argtypes = [i32, _f64p, _f64p, sz, _u8p]
    L.ocbh_fit_inliers.restype = None
    L.ocbh_check_sample_degeneracy_h.argtypes = [_f64p, sz, _szp]
    L.ocbh_check_degeneracy_f.argtypes = [_f64p, dbl, _f64p, sz, _u8p]
    L.ocbh_decompose_essential.argtypes = [_f64p, _f64p]
    L.ocbh_decompose_essential.restype = None
    L.ocbh_decompose_homography.argtypes = [_f64p, _f64p, sz, _u8p, _f64p]
    L.ocbh_decompose_homography.restype = i32
    L.ocbh_decompose_homography_mat.argtypes = [_f64p, _f64p, _f64p, _f64p]
    L.ocbh_decompose_homography_mat.restype = i32
    L.ocbh_assemble_inliers.argtypes = [_szp, _szp, _f64p, sz, _u8p, _f64p, sz, _f64p, sz, _f64p, _szp]
    L.ocbh_assemble_inliers.restype = sz
    L.ocbh_full_piv_lu_solve.argtypes = [_f64p, i32, i32, _f64p, _f64p]
    L.ocbh_full_piv_lu_solve.restype = None
    L.ocbh_invert3.argtypes = [_f64p, _f64p]
    L.ocbh_invert3.restype = None
    L.ocbh_jacobi_svd.argtypes = [_f64p, i32, _f64p, _f64p, _f64p]
    L.ocbh_jacobi_svd.restype = None
    L.ocbh_jacobi_svd_tall.argtypes = [_f64p, i32, i32, _f64p, _f64p]
    L.ocbh_jacobi_svd_tall.restype = None
    L.ocbh_run_parallel_match.argtypes = [_u64p, _u64p, sz, sz, sz, i32, _szp, _f64p]
    L.ocbh_image_to_3d.argtypes = [_f64p, sz, _f64p, _f64p]
    L.ocbh_image_to_3d.restype = None
    L.ocbh_link_pairs.argtypes = [C.c_void_p, _szp, _f64p, sz, _szp, sz, i32, sz, i32, dbl, i32, i32, vp, vp, sz, vp, vp]
    L.ocbh_link_pairs.restype = C.c_void_p
    L.ocbh_link_pack_matches.argtypes = [C.c_void_p, vp, vp, i32]
    L.ocbh_link_pack_matches.restype = sz
    L.ocbh_partition_pairs.argtypes = [vp, sz, _szp, sz, sz, vp, vp, vp]
    L.ocbh_prosac_order.argtypes = [_f64p, sz, vp]
    L.ocbh_prosac_order.restype = None
    L.ocbh_hilbert_index.argtypes = [i32, i32, i32]
    L.ocbh_hilbert_index.restype = C.c_uint32
    L.ocbh_hilbert_order.argtypes = [_f64p, sz, _szp]
    L.ocbh_hilbert_order.restype = None
    L.ocbh_link_free.argtypes = [C.c_void_p]
    L.ocbh_link_free.restype = None
    L.ocbh_link_stats.argtypes = [C.c_void_p, _f64p]
    L.ocbh_link_stats.restype = None
    L.ocbh_link_sizes.argtypes = [C.c_void_p, sz, _szp, _szp]
    L.ocbh_link_sizes.restype = None
    L.ocbh_link_get.argtypes = [C.c_void_p, sz, _szp, _szp, _f64p, _f64p, C.c_void_p, _f64p, _f64p, _szp]
    L.ocbh_link_get.restype = None
    L.ocbh_ransac_batch.argtypes = [i32, _f64p, _szp, sz, i32, _f64p, _f64p, _u8p, _szp]
    L.ocbh_refit_evaluate_batch.restype = i32
    L.ocbh_refit_evaluate_batch.argtypes = [_f64p, _szp, sz, i32, C.c_double, _f64p, _f64p, _u8p]
    L.ocbh_run_parallel_handles.argtypes = [C.c_void_p, C.c_void_p, sz, i32, i32, i32, _szp, _f64p]
    _lib = L
    return L


def _check(rc):
    if rc != 0:
        raise OcbError("host mirror: " + lib().ocbh_last_error().decode())


def _rows(a):
    a = np.ascontiguousarray(a)
    if a.dtype != np.uint64:
        a = a.view(np.uint64)
    return a.reshape(-1, 8)


def _corr(c):
    return np.ascontiguousarray(c, np.float64).reshape(-1, 7)


def layout():
    L = lib()
    return dict(sizeof_feature_2d=L.ocbh_sizeof_feature_2d(), offsetof_descriptor=L.ocbh_offsetof_descriptor(),
                sizeof_feature_match=L.ocbh_sizeof_feature_match(),
                sizeof_correspondence=L.ocbh_sizeof_correspondence(),
                sizeof_feature_match_denormalized=L.ocbh_sizeof_feature_match_denormalized())


# ---- src/match ----
def match_features_subset(desc1, desc2, idx1, idx2, cross_check=False):
    """-> (feature_index_1, feature_index_2, distance[, mutual]) in the reference's output order."""
    desc1, desc2 = _rows(desc1), _rows(desc2)
    idx1 = np.ascontiguousarray(idx1, np.uintp)
    idx2 = np.ascontiguousarray(idx2, np.uintp)
    n1 = len(idx1)
    o1, o2, od = np.zeros(max(n1, 1), np.uintp), np.zeros(max(n1, 1), np.uintp), np.zeros(max(n1, 1), np.float64)
    mut = np.zeros(max(n1, 1), np.uint8) if cross_check else None
    n = np.zeros(1, np.uintp)
    _check(lib().ocbh_match_features_subset(desc1, len(desc1), desc2, len(desc2), idx1, n1, idx2, len(idx2), o1, o2, od,
                                            None if mut is None else mut.ctypes.data_as(C.c_void_p), n))
    m = int(n[0])
    if cross_check:
        return o1[:m].copy(), o2[:m].copy(), od[:m].copy(), mut[:m].astype(bool)
    return o1[:m].copy(), o2[:m].copy(), od[:m].copy()


def match_features_guided(desc1, desc2, query_feature, begin, nearby):
    """Guided matcher of the dense stage (src/dense/dense_stereo.cpp:244-281) for lists in CSR form.
    -> (list, query_feature, candidate_feature, best_distance, second_distance) of the accepted visits, list order."""
    desc1, desc2 = _rows(desc1), _rows(desc2)
    query_feature = np.ascontiguousarray(query_feature, np.uintp)
    begin = np.ascontiguousarray(begin, np.uintp)
    nearby = np.ascontiguousarray(nearby, np.uintp)
    nl = len(query_feature)
    assert len(begin) == nl + 1
    m = max(nl, 1)
    ol, oq, oc = np.zeros(m, np.uintp), np.zeros(m, np.uintp), np.zeros(m, np.uintp)
    ob, os_ = np.zeros(m, np.float64), np.zeros(m, np.float64)
    n = np.zeros(1, np.uintp)
    if len(nearby) == 0:
        nearby = np.zeros(1, np.uintp)
    _check(lib().ocbh_match_guided(desc1, len(desc1), desc2, len(desc2), query_feature if nl else np.zeros(1, np.uintp),
                                   begin, nearby, nl, ol, oq, oc, ob, os_, n))
    k = int(n[0])
    return ol[:k].copy(), oq[:k].copy(), oc[:k].copy(), ob[:k].copy(), os_[:k].copy()


class FeatureSet:
    """A std::vector<feature_2d> living on the C++ side (what the pipeline holds per image)."""

    def __init__(self, desc, xy=None, strength=None):
        desc = _rows(desc)
        self.n = len(desc)
        xyp = None if xy is None else np.ascontiguousarray(xy, np.float64).ctypes.data_as(C.c_void_p)
        stp = None if strength is None else np.ascontiguousarray(strength, np.float32).ctypes.data_as(C.c_void_p)
        self.handle = lib().ocbh_features_create(xyp, stp, desc, self.n)

    def close(self):
        if self.handle:
            lib().ocbh_features_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


class Matcher:
    """match_features_subset(set_1, set_2, indices_1, indices_2) on two FeatureSets with preallocated outputs."""

    def __init__(self, max_queries):
        m = max(int(max_queries), 1)
        self.o1, self.o2 = np.zeros(m, np.uintp), np.zeros(m, np.uintp)
        self.od, self.mut, self.n = np.zeros(m, np.float64), np.zeros(m, np.uint8), np.zeros(1, np.uintp)

    def __call__(self, set_1, set_2, idx1, idx2, cross_check=False):
        _check(lib().ocbh_match_handles(set_1.handle, set_2.handle, idx1, len(idx1), idx2, len(idx2), self.o1, self.o2,
                                        self.od, self.mut.ctypes.data_as(C.c_void_p) if cross_check else None, self.n))
        return int(self.n[0])


def spatially_subsample_feature_indices(xy, strength, spacing, count=0):
    xy = np.ascontiguousarray(xy, np.float64).reshape(-1, 2)
    strength = np.ascontiguousarray(strength, np.float32)
    out = np.zeros(max(len(xy), 1), np.uintp)
    m = lib().ocbh_subsample(xy, strength, len(xy), float(spacing), int(count), out)
    return out[:m].copy()


# ---- src/model_inliers ----
def ransac(kind, corr):
    """-> (score, M18, inliers, stats)"""
    corr = _corr(corr)
    M18 = np.full(18, np.nan)
    inl = np.zeros(max(len(corr), 1), np.uint8)
    score = np.zeros(1)
    st = np.zeros(6, np.uintp)
    _check(lib().ocbh_ransac(kind, corr, len(corr), M18, inl, score, st))
    stats = dict(iterations=int(st[0]), improvements=int(st[1]), rejected=int(st[2]), degenerate=int(st[3]),
                 scored=int(st[4]), gpu_calls=int(st[5]))
    return float(score[0]), M18, inl[:len(corr)].astype(bool), stats


def set_ransac_device_fit(on):
    """ransac(homography): fit each batch of minimal samples on the device (default) or on the host."""
    lib().ocbh_set_ransac_device_fit(1 if on else 0)


def evaluate(kind, M18, corr, thr=0.0):
    corr = _corr(corr)
    inl = np.zeros(max(len(corr), 1), np.uint8)
    score = np.zeros(1)
    _check(lib().ocbh_evaluate(kind, np.ascontiguousarray(M18, np.float64), float(thr), corr, len(corr), inl, score))
    return float(score[0]), inl[:len(corr)].astype(bool)


def error(kind, M18, corr7):
    return lib().ocbh_error(kind, np.ascontiguousarray(M18, np.float64), np.ascontiguousarray(corr7, np.float64))


def fit(kind, corr, sample):
    corr = _corr(corr)
    M18 = np.full(18, np.nan)
    lib().ocbh_fit(kind, corr, len(corr), np.ascontiguousarray(sample, np.uintp), M18)
    return M18


def fit_inliers(kind, M18, corr, inliers):
    corr = _corr(corr)
    M18 = np.array(M18, np.float64)
    lib().ocbh_fit_inliers(kind, M18, corr, len(corr), np.ascontiguousarray(inliers, np.uint8))
    return M18


def check_sample_degeneracy_h(corr, sample):
    corr = _corr(corr)
    return bool(lib().ocbh_check_sample_degeneracy_h(corr, len(corr), np.ascontiguousarray(sample, np.uintp)))


def check_degeneracy_f(M18, corr, inliers, thr=0.01):
    corr = _corr(corr)
    M18 = np.array(M18, np.float64)
    inl = np.ascontiguousarray(inliers, np.uint8).copy()
    _check(lib().ocbh_check_degeneracy_f(M18, float(thr), corr, len(corr), inl))
    return M18, inl.astype(bool)


def decompose_essential(M18):
    out = np.zeros(28)
    lib().ocbh_decompose_essential(np.ascontiguousarray(M18, np.float64), out)
    return out.reshape(4, 7)


def decompose_homography(M18, corr, inliers):
    """homography_model::decompose -> (ok, poses [4][8] = qx,qy,qz,qw, tx,ty,tz, score) in the reference's order."""
    corr = _corr(corr)
    out = np.full(32, np.nan)
    inl = np.ascontiguousarray(inliers, np.uint8)
    ok = lib().ocbh_decompose_homography(np.ascontiguousarray(M18, np.float64), corr, len(corr), inl, out)
    return bool(ok), out.reshape(4, 8)


def decompose_homography_mat(H):
    """cv::decomposeHomographyMat(H, I) restated: -> (Rs [k][3][3], ts [k][3], ns [k][3])"""
    H9 = np.ascontiguousarray(np.asarray(H, np.float64).T).ravel()
    R, t, n = np.zeros(36), np.zeros(12), np.zeros(12)
    k = lib().ocbh_decompose_homography_mat(H9, R, t, n)
    return R.reshape(4, 3, 3).transpose(0, 2, 1)[:k].copy(), t.reshape(4, 3)[:k].copy(), n.reshape(4, 3)[:k].copy()


def assemble_inliers(m_i1, m_i2, m_dist, inliers, xy1, xy2):
    m_i1 = np.ascontiguousarray(m_i1, np.uintp)
    m_i2 = np.ascontiguousarray(m_i2, np.uintp)
    m_dist = np.ascontiguousarray(m_dist, np.float64)
    xy1 = np.ascontiguousarray(xy1, np.float64).reshape(-1, 2)
    xy2 = np.ascontiguousarray(xy2, np.float64).reshape(-1, 2)
    n = len(m_i1)
    px, ix = np.zeros((max(n, 1), 4)), np.zeros((max(n, 1), 3), np.uintp)
    m = lib().ocbh_assemble_inliers(m_i1, m_i2, m_dist, n, np.ascontiguousarray(inliers, np.uint8), xy1, len(xy1), xy2,
                                    len(xy2), px, ix)
    return px[:m].copy(), ix[:m].copy()


# ---- host linear algebra (numpy in/out, row-major views) ----
def full_piv_lu_solve(A, b):
    A = np.asarray(A, np.float64)
    x = np.zeros(A.shape[1])
    lib().ocbh_full_piv_lu_solve(np.ascontiguousarray(A.T).ravel(), A.shape[0], A.shape[1],
                                 np.ascontiguousarray(b, np.float64), x)
    return x


def invert3(M):
    out = np.zeros(9)
    lib().ocbh_invert3(np.ascontiguousarray(np.asarray(M, np.float64).T).ravel(), out)
    return out.reshape(3, 3).T.copy()


def jacobi_svd(A):
    A = np.asarray(A, np.float64)
    n = A.shape[0]
    U, S, V = np.zeros(n * n), np.zeros(n), np.zeros(n * n)
    lib().ocbh_jacobi_svd(np.ascontiguousarray(A.T).ravel(), n, U, S, V)
    return U.reshape(n, n).T.copy(), S, V.reshape(n, n).T.copy()


def jacobi_svd_tall(A):
    A = np.asarray(A, np.float64)
    r, c = A.shape
    S, V = np.zeros(c), np.zeros(c * c)
    lib().ocbh_jacobi_svd_tall(np.ascontiguousarray(A.T).ravel(), r, c, S, V)
    return S, V.reshape(c, c).T.copy()


def run_parallel_match(q, c, n_pairs, n1, n2, threads=0):
    """The reference's run_parallel shape: one match_features_subset closure per pair on `threads` OpenMP workers."""
    q, c = _rows(q), _rows(c)
    nm, secs = np.zeros(1, np.uintp), np.zeros(1)
    _check(lib().ocbh_run_parallel_match(q, c, n_pairs, n1, n2, threads, nm, secs))
    return float(secs[0]), int(nm[0])


def run_parallel_handles(sets_q, sets_c, threads=0, cross_check=False, reps=1):
    """reps * len(sets_q) match_features_subset closures over FeatureSets on `threads` OpenMP workers ->
    (wall seconds, total matches)."""
    n = len(sets_q)
    hq = (C.c_void_p * n)(*[s.handle for s in sets_q])
    hc = (C.c_void_p * n)(*[s.handle for s in sets_c])
    nm, secs = np.zeros(1, np.uintp), np.zeros(1)
    _check(lib().ocbh_run_parallel_handles(hq, hc, n, int(threads), int(cross_check), int(reps), nm, secs))
    return float(secs[0]), int(nm[0])


# ---- src/distort (host) ----
def camera8(f, pp, radial=(0, 0, 0), tangential=(0, 0)):
    """(f, ppx, ppy, k1, k2, k3, p1, p2): the DifferentiableCameraModel<double> members image_to_3d reads."""
    return np.array([f, pp[0], pp[1], *radial, *tangential], np.float64)


def image_to_3d(xy, cam8):
    xy = np.ascontiguousarray(xy, np.float64).reshape(-1, 2)
    rays = np.zeros((len(xy), 3))
    lib().ocbh_image_to_3d(xy, len(xy), np.ascontiguousarray(cam8, np.float64), rays)
    return rays


# ---- batched LinkStage runner (host/link_batch.hpp) ----
class LinkResults:
    """Per-pair camera_relations of one link_pairs call (kept on the C++ side; get(p) flattens one pair)."""

    def __init__(self, handle, n_pairs):
        self.handle, self.n_pairs = handle, n_pairs
        st = np.zeros(9)
        lib().ocbh_link_stats(handle, st)
        self.stats = dict(seconds_subsample_upload=st[0], seconds_match_gpu=st[1], seconds_tail=st[2],
                          seconds_total=st[3], seconds_setup=st[7], seconds_release=st[8], comparisons=int(st[4]), matches=int(st[5]), ransac_inliers=int(st[6]))

    def sizes(self, p):
        a, b = np.zeros(1, np.uintp), np.zeros(1, np.uintp)
        lib().ocbh_link_sizes(self.handle, p, a, b)
        return int(a[0]), int(b[0])

    def get(self, p):
        nm, ni = self.sizes(p)
        i1, i2, d = np.zeros(max(nm, 1), np.uintp), np.zeros(max(nm, 1), np.uintp), np.zeros(max(nm, 1))
        H9, poses, rt = np.zeros(9), np.zeros(32), C.c_int(0)
        px, ix = np.zeros((max(ni, 1), 4)), np.zeros((max(ni, 1), 3), np.uintp)
        lib().ocbh_link_get(self.handle, p, i1, i2, d, H9, C.byref(rt), poses, px, ix)
        return dict(matches=(i1[:nm].copy(), i2[:nm].copy(), d[:nm].copy()), H=H9.reshape(3, 3).T.copy(),
                    relation_type=rt.value, poses=poses.reshape(4, 8).copy(), inlier_pixels=px[:ni].copy(),
                    inlier_idx=ix[:ni].copy())

    def pack_matches(self, out=None, threads=0):
        """All match lists as flat records -> (counts [n_pairs] uint64, records [total][3] uint32 = feature_index_1,
        feature_index_2, integer Hamming distance), pair after pair. `out`: optional preallocated uint32 buffer (e.g. a
        view of shared memory) that receives the records."""
        counts = np.zeros(max(self.n_pairs, 1), np.uint64)
        total = int(lib().ocbh_link_pack_matches(self.handle, counts.ctypes.data_as(C.c_void_p), None, 0))
        if out is None:
            out = np.zeros(max(total, 1) * 3, np.uint32)
        assert out.dtype == np.uint32 and out.size >= total * 3 and out.flags["C_CONTIGUOUS"]
        lib().ocbh_link_pack_matches(self.handle, counts.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p),
                                     int(threads))
        return counts[:self.n_pairs], out[:total * 3].reshape(total, 3)

    def close(self):
        if self.handle:
            lib().ocbh_link_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


def link_pairs(feature_sets, cameras8, pairs, num_sparse=None, threads=0, pairs_per_submission=0, run_ransac=True,
               spacing=0.0, device_tail=True, n_devices=1, positions=None, packed=None):
    """What LinkStage's closures compute for every (image, neighbour) pair (src/pipeline/link_stage.cpp:75-112), for
    the whole pair list: feature_sets = FeatureSet per image, cameras8 = camera8() per image, pairs = [(i, j)].
    device_tail=False keeps the ratio test and the rays on the host (A/B). n_devices > 1: the pair list is partitioned
    over that many GPUs of this process along the Hilbert curve of `positions` ([n][2]; None: image index order).
    packed = (records uint32 flat buffer, offsets uint64 [n_pairs], counts uint64 [n_pairs]): the tail workers also write
    every pair's final match list as 12-byte records into `records` (LinkOptions::packed_out)."""
    n = len(feature_sets)
    h = (C.c_void_p * n)(*[s.handle for s in feature_sets])
    ns = np.zeros(n, np.uintp) if num_sparse is None else np.ascontiguousarray(num_sparse, np.uintp)
    cams = np.ascontiguousarray(cameras8, np.float64).reshape(n, 8)
    pr = np.ascontiguousarray(pairs, np.uintp).reshape(-1, 2)
    pos = None if positions is None else np.ascontiguousarray(positions, np.float64).reshape(n, 2)
    res = lib().ocbh_link_pairs(h, ns, cams, n, pr, len(pr), int(threads), int(pairs_per_submission), int(run_ransac),
                               float(spacing), int(bool(device_tail)), int(n_devices),
                               None if pos is None else pos.ctypes.data_as(C.c_void_p),
                               *((None, 0, None, None) if packed is None else
                                 (packed[0].ctypes.data_as(C.c_void_p), packed[0].size // 3,
                                  packed[1].ctypes.data_as(C.c_void_p), packed[2].ctypes.data_as(C.c_void_p))))
    if not res:
        raise OcbError("host mirror: " + lib().ocbh_last_error().decode())
    return LinkResults(res, len(pr))


def refit_evaluate_batch(corr_list, inlier_list, rounds=3, thr=0.0):
    """The loop of RelaxGroup::finalize (src/relax/relax_group.cpp:156-165) for many edges in lock step:
    rounds x (homography fitInliers, evaluate) -> list of (score, M18, inliers)."""
    corr_list = [_corr(c) for c in corr_list]
    n = len(corr_list)
    offsets = np.zeros(n + 1, np.uintp)
    offsets[1:] = np.cumsum([len(c) for c in corr_list])
    total = int(offsets[-1])
    allc = np.concatenate(corr_list) if total else np.zeros((0, 7))
    inl = np.zeros(max(total, 1), np.uint8)
    for j, f in enumerate(inlier_list):
        inl[int(offsets[j]):int(offsets[j + 1])] = np.asarray(f, np.uint8)
    scores, M18 = np.zeros(max(n, 1)), np.full((max(n, 1), 18), np.nan)
    _check(lib().ocbh_refit_evaluate_batch(np.ascontiguousarray(allc), offsets, n, int(rounds), float(thr), scores, M18, inl))
    return [(float(scores[j]), M18[j].copy(), inl[int(offsets[j]):int(offsets[j + 1])].astype(bool)) for j in range(n)]


def ransac_batch(kind, corr_list, threads=0):
    """ransac<Model>() for every correspondence set of corr_list, advanced in lock step (one GPU launch per round)
    -> list of (score, M18, inliers, stats)."""
    corr_list = [_corr(c) for c in corr_list]
    n = len(corr_list)
    offsets = np.zeros(n + 1, np.uintp)
    offsets[1:] = np.cumsum([len(c) for c in corr_list])
    total = int(offsets[-1])
    allc = np.concatenate(corr_list) if total else np.zeros((0, 7))
    scores, M18 = np.zeros(max(n, 1)), np.full((max(n, 1), 18), np.nan)
    inl, st = np.zeros(max(total, 1), np.uint8), np.zeros((max(n, 1), 2), np.uintp)
    _check(lib().ocbh_ransac_batch(kind, np.ascontiguousarray(allc), offsets, n, int(threads), scores, M18, inl, st))
    return [(float(scores[j]), M18[j].copy(), inl[int(offsets[j]):int(offsets[j + 1])].astype(bool),
             dict(iterations=int(st[j, 0]), improvements=int(st[j, 1]))) for j in range(n)]


# ---- overlap-graph partition (host/partition.hpp) ----
def hilbert_index(order, x, y):
    return int(lib().ocbh_hilbert_index(int(order), int(x), int(y)))


def hilbert_order(positions):
    pos = np.ascontiguousarray(positions, np.float64).reshape(-1, 2)
    out = np.zeros(max(len(pos), 1), np.uintp)
    if len(pos):
        lib().ocbh_hilbert_order(pos, len(pos), out)
    return out[:len(pos)].astype(np.int64)


def partition_pairs(positions, pairs, world):
    """-> (owner [n_images] part of every image, pair_part [n_pairs] part of every pair, halo [world][n_images] bool).
    positions None: the images' index order stands in for the Hilbert curve."""
    pr = np.ascontiguousarray(pairs, np.uintp).reshape(-1, 2)
    if positions is None:
        raise ValueError("positions are required (pass an [n][2] array)")
    pos = np.ascontiguousarray(positions, np.float64).reshape(-1, 2)
    n = len(pos)
    owner, part = np.zeros(max(n, 1), np.uint32), np.zeros(max(len(pr), 1), np.uint32)
    halo = np.zeros((world, max(n, 1)), np.uint8)
    _check(lib().ocbh_partition_pairs(pos.ctypes.data_as(C.c_void_p), n, pr, len(pr), int(world),
                                      owner.ctypes.data_as(C.c_void_p), part.ctypes.data_as(C.c_void_p),
                                      halo.ctypes.data_as(C.c_void_p)))
    return owner[:n], part[:len(pr)], halo[:, :n].astype(bool)


def prosac_order(quality):
    """ransac.cpp:83-90 on the host: indices 0 .. n-1 sorted by quality, ascending, with libstdc++'s std::sort."""
    q = np.ascontiguousarray(quality, np.float64)
    out = np.zeros(max(len(q), 1), np.uint32)
    lib().ocbh_prosac_order(q, len(q), out.ctypes.data_as(C.c_void_p))
    return out[:len(q)]
