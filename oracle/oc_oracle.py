"""TEST INFRASTRUCTURE ONLY -- ctypes bindings for the CPU oracle (oracle/liboc_oracle.so, the
restatement) and, when it has been built, for the reference's own object code (oracle/_ref/liboc_ref*.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
The product package (opencalibration_b200) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = "/root/reference"

KIND_H, KIND_E, KIND_F = 0, 1, 2
MIN_POINTS = {KIND_H: 4, KIND_E: 5, KIND_F: 8}
DEFAULT_THR = {KIND_H: 0.005, KIND_E: 0.01, KIND_F: 0.01}

_u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
_u16p = np.ctypeslib.ndpointer(np.uint16, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_szp = np.ctypeslib.ndpointer(np.uintp, flags="C_CONTIGUOUS")


def build(ref=True, quiet=True):
    """Compile the oracle (and oracle/_ref when /root/reference exists). Building the checker is not using it."""
    targets = ["port"]
    if ref and os.path.isdir(REF_ROOT):
        targets.append("ref")
    subprocess.run(["make", "-C", HERE, "-j4"] + targets, check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def _rows(a):
    a = np.ascontiguousarray(a)
    if a.dtype != np.uint64:
        a = a.view(np.uint64)
    return a.reshape(-1, 8)


def _opt(arr):
    return None if arr is None else arr.ctypes.data_as(C.c_void_p)


class Oracle:
    """The restatement ("port")."""

    def __init__(self, path=None):
        path = path or os.path.join(HERE, "liboc_oracle.so")
        if not os.path.exists(path):
            build(ref=False)
        L = self.lib = C.CDLL(path)
        L.oco_match_top2.argtypes = [_u64p, C.c_size_t, _u64p, C.c_size_t, _u32p, _u16p, _u16p]
        L.oco_match_col_best.argtypes = [_u64p, C.c_size_t, _u64p, C.c_size_t, _u32p]
        L.oco_match_features_subset.argtypes = [_u64p, _u64p, _szp, C.c_size_t, _szp, C.c_size_t, _szp, _szp, _f64p]
        L.oco_match_features_subset.restype = C.c_size_t
        L.oco_subsample.argtypes = [_f64p, _f32p, C.c_size_t, C.c_double, C.c_size_t, _szp]
        L.oco_subsample.restype = C.c_size_t
        L.oco_error.argtypes = [C.c_int, _f64p, C.c_double, _f64p]
        L.oco_error.restype = C.c_double
        L.oco_evaluate.argtypes = [C.c_int, _f64p, C.c_double, _f64p, C.c_size_t, _u8p]
        L.oco_evaluate.restype = C.c_double
        L.oco_fit.argtypes = [C.c_int, _f64p, _szp, _f64p]
        L.oco_fit_inliers.argtypes = [C.c_int, _f64p, _f64p, C.c_size_t, _u8p]
        L.oco_check_sample_degeneracy_h.argtypes = [_f64p, _szp]
        L.oco_check_sample_degeneracy_h.restype = C.c_int
        L.oco_check_degeneracy_f.argtypes = [_f64p, C.c_double, _f64p, C.c_size_t, _u8p]
        L.oco_ransac.argtypes = [C.c_int, _f64p, C.c_size_t, _f64p, _u8p, _szp]
        L.oco_ransac.restype = C.c_double
        L.oco_hypothesis_stream.argtypes = [C.c_int, _f64p, C.c_size_t, C.c_size_t, _szp, _szp]
        L.oco_hypothesis_stream.restype = C.c_int
        L.oco_score_hypotheses.argtypes = [C.c_int, _f64p, C.c_size_t, C.c_double, _f64p, C.c_size_t, C.c_void_p,
                                           _f64p, _u32p, C.c_void_p, C.c_int]
        L.oco_fullpivlu_solve.argtypes = [_f64p, C.c_int, C.c_int, _f64p, _f64p]
        L.oco_inverse3.argtypes = [_f64p, _f64p]
        L.oco_jacobi_svd_square.argtypes = [_f64p, C.c_int, C.c_void_p, _f64p, C.c_void_p]
        L.oco_jacobi_svd_tall_v.argtypes = [_f64p, C.c_int, C.c_int, _f64p, _f64p]
        L.oco_bench_match_pairs.argtypes = [_u64p, _u64p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, _szp]
        L.oco_bench_match_pairs.restype = C.c_double
        L.oco_num_procs.restype = C.c_int
        L.oco_match_lists.argtypes = [_u64p, _u64p, _u32p, _u64p, _u32p, C.c_size_t, _u32p, _f64p, _f64p, _u8p]
        L.oco_bench_match_lists.argtypes = [_u64p, _u64p, _u32p, _u64p, _u32p, C.c_size_t, C.c_int, _u32p, _f64p, _f64p]
        L.oco_bench_match_lists.restype = C.c_double
        L.oco_scene_homography.argtypes = [C.c_size_t, C.c_size_t, C.c_uint, _f64p, _f64p]
        L.oco_scene_homography_near_degenerate.argtypes = [_f64p, _f64p]
        L.oco_scene_fundamental.argtypes = [C.c_size_t, C.c_size_t, C.c_double, C.c_uint, _f64p, _f64p]

    # ---- synthetic scenes of test/test_ransac_benchmark.cpp (returns corr [n][7], ground truth 3x3) ----
    def scene_homography(self, n_inliers, n_outliers, seed=42):
        corr, gt = np.zeros((n_inliers + n_outliers, 7)), np.zeros(9)
        self.lib.oco_scene_homography(n_inliers, n_outliers, seed, corr, gt)
        return corr, gt.reshape(3, 3)

    def scene_homography_near_degenerate(self):
        corr, gt = np.zeros((100, 7)), np.zeros(9)
        self.lib.oco_scene_homography_near_degenerate(corr, gt)
        return corr, gt.reshape(3, 3)

    def scene_fundamental(self, n_inliers, n_outliers, planar_fraction=0.0, seed=42):
        corr, gt = np.zeros((n_inliers + n_outliers, 7)), np.zeros(9)
        self.lib.oco_scene_fundamental(n_inliers, n_outliers, planar_fraction, seed, corr, gt)
        return corr, gt.reshape(3, 3)

    # ---- match ----
    def match_top2(self, q, c):
        q, c = _rows(q), _rows(c)
        n1 = len(q)
        bk, bd, sd = np.zeros(n1, np.uint32), np.zeros(n1, np.uint16), np.zeros(n1, np.uint16)
        self.lib.oco_match_top2(q, n1, c, len(c), bk, bd, sd)
        return bk, bd, sd

    def match_lists(self, q, c, list_query, list_begin, list_candidates):
        """src/dense/dense_stereo.cpp:251-276 -> (best position, best distance, second distance, accepted) per list."""
        q, c = _rows(q), _rows(c)
        lq = np.ascontiguousarray(list_query, np.uint32)
        lb = np.ascontiguousarray(list_begin, np.uint64)
        lc = np.ascontiguousarray(list_candidates, np.uint32)
        if len(lc) == 0:
            lc = np.zeros(1, np.uint32)
        if len(c) == 0:
            c = np.zeros((1, 8), np.uint64)
        nl = len(lq)
        bp, bd, sd, good = np.zeros(nl, np.uint32), np.zeros(nl), np.zeros(nl), np.zeros(nl, np.uint8)
        if nl:
            self.lib.oco_match_lists(q, c, lq, lb, lc, nl, bp, bd, sd, good)
        return bp, bd, sd, good.astype(bool)

    def bench_match_lists(self, q, c, list_query, list_begin, list_candidates, threads=0):
        q, c = _rows(q), _rows(c)
        lq = np.ascontiguousarray(list_query, np.uint32)
        lb = np.ascontiguousarray(list_begin, np.uint64)
        lc = np.ascontiguousarray(list_candidates, np.uint32)
        nl = len(lq)
        bp, bd, sd = np.zeros(nl, np.uint32), np.zeros(nl), np.zeros(nl)
        return self.lib.oco_bench_match_lists(q, c, lq, lb, lc, nl, threads, bp, bd, sd)

    def match_col_best(self, q, c):
        q, c = _rows(q), _rows(c)
        out = np.zeros(len(c), np.uint32)
        self.lib.oco_match_col_best(q, len(q), c, len(c), out)
        return out

    def match_features_subset(self, desc1, desc2, idx1, idx2):
        desc1, desc2 = _rows(desc1), _rows(desc2)
        idx1 = np.ascontiguousarray(idx1, np.uintp)
        idx2 = np.ascontiguousarray(idx2, np.uintp)
        n1 = len(idx1)
        o1, o2, od = np.zeros(n1, np.uintp), np.zeros(n1, np.uintp), np.zeros(n1, np.float64)
        m = self.lib.oco_match_features_subset(desc1, desc2, idx1, n1, idx2, len(idx2), o1, o2, od)
        return o1[:m].copy(), o2[:m].copy(), od[:m].copy()

    def subsample(self, xy, strength, spacing, count=0):
        xy = np.ascontiguousarray(xy, np.float64).reshape(-1, 2)
        strength = np.ascontiguousarray(strength, np.float32)
        out = np.zeros(max(len(xy), 1), np.uintp)
        m = self.lib.oco_subsample(xy, strength, len(xy), float(spacing), int(count), out)
        return out[:m].copy()

    # ---- models ----
    @staticmethod
    def _corr(corr):
        return np.ascontiguousarray(corr, np.float64).reshape(-1, 7)

    def error(self, kind, M18, corr7, thr=0.0):
        return self.lib.oco_error(kind, np.ascontiguousarray(M18, np.float64), thr,
                                  np.ascontiguousarray(corr7, np.float64))

    def evaluate(self, kind, M18, corr, thr=0.0):
        corr = self._corr(corr)
        inl = np.zeros(len(corr), np.uint8)
        s = self.lib.oco_evaluate(kind, np.ascontiguousarray(M18, np.float64), thr, corr, len(corr), inl)
        return s, inl.astype(bool)

    def fit(self, kind, corr, sample):
        M18 = np.full(18, np.nan)
        self.lib.oco_fit(kind, self._corr(corr), np.ascontiguousarray(sample, np.uintp), M18)
        return M18

    def fit_inliers(self, kind, M18, corr, inliers):
        corr = self._corr(corr)
        M18 = np.array(M18, np.float64)
        self.lib.oco_fit_inliers(kind, M18, corr, len(corr), np.ascontiguousarray(inliers, np.uint8))
        return M18

    def check_sample_degeneracy_h(self, corr, sample):
        return bool(self.lib.oco_check_sample_degeneracy_h(self._corr(corr), np.ascontiguousarray(sample, np.uintp)))

    def check_degeneracy_f(self, M18, corr, inliers, thr=0.01):
        corr = self._corr(corr)
        M18 = np.array(M18, np.float64)
        inl = np.ascontiguousarray(inliers, np.uint8).copy()
        self.lib.oco_check_degeneracy_f(M18, thr, corr, len(corr), inl)
        return M18, inl.astype(bool)

    def ransac(self, kind, corr):
        corr = self._corr(corr)
        M18 = np.full(18, np.nan)
        inl = np.zeros(max(len(corr), 1), np.uint8)
        tr = np.zeros(4, np.uintp)
        s = self.lib.oco_ransac(kind, corr, len(corr), M18, inl, tr)
        return s, M18, inl[:len(corr)].astype(bool), dict(iterations=int(tr[0]), improvements=int(tr[1]),
                                                          rejected=int(tr[2]), degenerate=int(tr[3]))

    def hypothesis_stream(self, kind, corr, count):
        corr = self._corr(corr)
        n = len(corr)
        eo = np.zeros(max(n, 1), np.uintp)
        sm = np.zeros(max(count * MIN_POINTS[kind], 1), np.uintp)
        ok = self.lib.oco_hypothesis_stream(kind, corr, n, count, eo, sm)
        if not ok:
            return None, None
        return eo[:n].copy(), sm[:count * MIN_POINTS[kind]].reshape(count, MIN_POINTS[kind]).copy()

    def score_hypotheses(self, kind, M18s, corr, order=None, thr=0.0, want_bits=True, threads=0):
        corr = self._corr(corr)
        M18s = np.ascontiguousarray(M18s, np.float64).reshape(-1, 18)
        h, n = len(M18s), len(corr)
        score, count = np.zeros(h, np.float64), np.zeros(h, np.uint32)
        bits = np.zeros((h, (n + 31) // 32), np.uint32) if want_bits else None
        order = None if order is None else np.ascontiguousarray(order, np.uintp)
        self.lib.oco_score_hypotheses(kind, M18s, h, thr, corr, n, _opt(order), score, count, _opt(bits), threads)
        return score, count, bits

    # ---- linear algebra (column-major in, numpy out) ----
    def fullpivlu_solve(self, A, b):
        A = np.asarray(A, np.float64)
        rows, cols = A.shape
        x = np.zeros(cols)
        self.lib.oco_fullpivlu_solve(np.asfortranarray(A).ravel(order="K").copy(), rows, cols,
                                     np.ascontiguousarray(b, np.float64), x)
        return x

    def inverse3(self, M):
        out = np.zeros(9)
        self.lib.oco_inverse3(np.asfortranarray(np.asarray(M, np.float64)).ravel(order="K").copy(), out)
        return out.reshape(3, 3).T.copy()

    def jacobi_svd_square(self, A):
        A = np.asarray(A, np.float64)
        n = A.shape[0]
        U, S, V = np.zeros(n * n), np.zeros(n), np.zeros(n * n)
        self.lib.oco_jacobi_svd_square(np.asfortranarray(A).ravel(order="K").copy(), n, _opt(U), S, _opt(V))
        return U.reshape(n, n).T.copy(), S, V.reshape(n, n).T.copy()

    def jacobi_svd_tall_v(self, A):
        A = np.asarray(A, np.float64)
        rows, cols = A.shape
        S, V = np.zeros(cols), np.zeros(cols * cols)
        self.lib.oco_jacobi_svd_tall_v(np.asfortranarray(A).ravel(order="K").copy(), rows, cols, S, V)
        return S, V.reshape(cols, cols).T.copy()

    # ---- CPU baseline timing ----
    def num_procs(self):
        return int(self.lib.oco_num_procs())

    def bench_match_pairs(self, q, c, n_pairs, n1, n2, threads=0):
        q, c = _rows(q), _rows(c)
        assert len(q) == n_pairs * n1 and len(c) == n_pairs * n2
        nm = np.zeros(1, np.uintp)
        secs = self.lib.oco_bench_match_pairs(q, c, n_pairs, n1, n2, threads, nm)
        return secs, int(nm[0])


class Reference:
    """The reference's own match_features.cpp / ransac.cpp object code (oracle/_ref)."""

    def __init__(self, popcnt=False):
        name = "liboc_ref_popcnt.so" if popcnt else "liboc_ref.so"
        path = os.path.join(HERE, "_ref", name)
        if not os.path.exists(path):
            if os.path.isdir(REF_ROOT):
                build(ref=True)
            else:
                raise FileNotFoundError(path)
        self.popcnt = popcnt
        L = self.lib = C.CDLL(path)
        for f in ("ocr_sizeof_feature_2d", "ocr_offsetof_descriptor", "ocr_sizeof_feature_match",
                  "ocr_sizeof_correspondence"):
            getattr(L, f).restype = C.c_size_t
        L.ocr_match_features_subset.argtypes = [_u64p, C.c_size_t, _u64p, C.c_size_t, _szp, C.c_size_t, _szp,
                                                C.c_size_t, _szp, _szp, _f64p]
        L.ocr_match_features_subset.restype = C.c_size_t
        L.ocr_subsample.argtypes = [_f64p, _f32p, C.c_size_t, C.c_double, C.c_size_t, _szp]
        L.ocr_subsample.restype = C.c_size_t
        L.ocr_ransac.argtypes = [C.c_int, _f64p, C.c_size_t, _f64p, _u8p]
        L.ocr_ransac.restype = C.c_double
        L.ocr_bench_match_pairs.argtypes = [_u64p, _u64p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, _szp]
        L.ocr_bench_match_pairs.restype = C.c_double
        L.ocr_num_procs.restype = C.c_int
        L.ocr_radius_lists.argtypes = [_f64p, C.c_size_t, _f64p, C.c_size_t, C.c_double, C.c_void_p, C.c_void_p]
        L.ocr_radius_lists.restype = C.c_size_t

    @staticmethod
    def available():
        return os.path.exists(os.path.join(HERE, "_ref", "liboc_ref.so")) or os.path.isdir(REF_ROOT)

    def layout(self):
        L = self.lib
        return dict(sizeof_feature_2d=L.ocr_sizeof_feature_2d(), offsetof_descriptor=L.ocr_offsetof_descriptor(),
                    sizeof_feature_match=L.ocr_sizeof_feature_match(),
                    sizeof_correspondence=L.ocr_sizeof_correspondence())

    def match_features_subset(self, desc1, desc2, idx1, idx2):
        desc1, desc2 = _rows(desc1), _rows(desc2)
        idx1 = np.ascontiguousarray(idx1, np.uintp)
        idx2 = np.ascontiguousarray(idx2, np.uintp)
        n1 = len(idx1)
        o1, o2, od = np.zeros(n1, np.uintp), np.zeros(n1, np.uintp), np.zeros(n1, np.float64)
        m = self.lib.ocr_match_features_subset(desc1, len(desc1), desc2, len(desc2), idx1, n1, idx2, len(idx2),
                                               o1, o2, od)
        return o1[:m].copy(), o2[:m].copy(), od[:m].copy()

    def radius_lists(self, cand_xy, pred_xy, radius=150.0):
        """Candidate lists of the dense stage from the reference's own jk-tree (dense_stereo.cpp:127-131,244-246):
        -> (begin [n_q + 1] uint64, nearby uint32) in the KD-tree's result order."""
        cand_xy = np.ascontiguousarray(cand_xy, np.float64).reshape(-1, 2)
        pred_xy = np.ascontiguousarray(pred_xy, np.float64).reshape(-1, 2)
        r2 = float(radius) * float(radius)
        total = self.lib.ocr_radius_lists(cand_xy, len(cand_xy), pred_xy, len(pred_xy), r2, None, None)
        begin, nearby = np.zeros(len(pred_xy) + 1, np.uint64), np.zeros(max(total, 1), np.uint32)
        self.lib.ocr_radius_lists(cand_xy, len(cand_xy), pred_xy, len(pred_xy), r2, begin.ctypes.data_as(C.c_void_p),
                                  nearby.ctypes.data_as(C.c_void_p))
        return begin, nearby[:total]

    def subsample(self, xy, strength, spacing, count=0):
        xy = np.ascontiguousarray(xy, np.float64).reshape(-1, 2)
        strength = np.ascontiguousarray(strength, np.float32)
        out = np.zeros(max(len(xy), 1), np.uintp)
        m = self.lib.ocr_subsample(xy, strength, len(xy), float(spacing), int(count), out)
        return out[:m].copy()

    def ransac(self, kind, corr):
        corr = np.ascontiguousarray(corr, np.float64).reshape(-1, 7)
        M18 = np.full(18, np.nan)
        inl = np.zeros(max(len(corr), 1), np.uint8)
        s = self.lib.ocr_ransac(kind, corr, len(corr), M18, inl)
        return s, M18, inl[:len(corr)].astype(bool)

    def num_procs(self):
        return int(self.lib.ocr_num_procs())

    def bench_match_pairs(self, q, c, n_pairs, n1, n2, threads=0):
        q, c = _rows(q), _rows(c)
        assert len(q) == n_pairs * n1 and len(c) == n_pairs * n2
        nm = np.zeros(1, np.uintp)
        secs = self.lib.ocr_bench_match_pairs(q, c, n_pairs, n1, n2, threads, nm)
        return secs, int(nm[0])
