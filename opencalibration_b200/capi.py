"""ctypes binding of include/ocb.h (libocb.so). The product path: raises if the CUDA library is missing."""
import ctypes as C
import os

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))

TOP2_DTYPE = np.dtype([("best_k", np.uint32), ("best_d", np.uint16), ("second_d", np.uint16)])
SET_SOURCE_DTYPE = np.dtype([("set_id", "<u8"), ("rows", "<u8"), ("stride", "<u8"), ("idx", "<u8"), ("n", "<u8")])
PAIR_DTYPE = np.dtype([("query_set", np.uint64), ("candidate_set", np.uint64)])
DIST_INF = 0xFFFF
NO_INDEX = 0xFFFFFFFF
MODEL_HOMOGRAPHY, MODEL_ESSENTIAL, MODEL_FUNDAMENTAL = 0, 1, 2


class OcbError(RuntimeError):
    pass


def lib_path():
    return os.path.join(PKG, "libocb.so")


_lib = None


def lib():
    """Load libocb.so (built in-tree by opencalibration_b200.build). No fallback: missing library is an error."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise OcbError(f"{path} not found: run `python -m opencalibration_b200.build` (there is no CPU fallback)")
    L = C.CDLL(path)
    vp, sz, i32, u64, dbl = C.c_void_p, C.c_size_t, C.c_int, C.c_uint64, C.c_double
    sigs = {
        "ocb_device_count": (i32, []),
        "ocb_init": (i32, [i32]),
        "ocb_set_device": (i32, [i32]),
        "ocb_current_device": (i32, []),
        "ocb_set_thread_blocking_sync": (i32, [i32]),
        "ocb_shutdown": (None, []),
        "ocb_last_error": (C.c_char_p, []),
        "ocb_version": (C.c_char_p, []),
        "ocb_kernel_launches": (u64, []),
        "ocb_set_option": (i32, [C.c_char_p, C.c_int64]),
        "ocb_get_option": (C.c_int64, [C.c_char_p]),
        "ocb_match_top2": (i32, [vp, sz, vp, sz, vp, vp]),
        "ocb_match_top2_strided": (i32, [vp, sz, vp, sz, vp, sz, vp, sz, vp, vp]),
        "ocb_match_top2_workspace_bytes": (sz, [sz, sz, i32]),
        "ocb_match_top2_device": (i32, [vp, sz, vp, sz, vp, vp, vp, sz, vp]),
        "ocb_register_descriptors": (i32, [u64, vp, sz]),
        "ocb_unregister_descriptors": (i32, [u64]),
        "ocb_register_descriptors_batch": (i32, [vp, sz]),
        "ocb_host_alloc": (vp, [sz]),
        "ocb_host_free": (None, [vp]),
        "ocb_match_pairs": (i32, [vp, sz, vp, vp]),
        "ocb_match_pairs_ratio": (i32, [vp, sz, vp, sz, vp]),
        "ocb_match_pairs_sorted": (i32, [vp, sz, vp, sz, vp, vp]),
        "ocb_register_images_batch": (i32, [vp, sz]),
        "ocb_corr_bind_batch_matches": (i32, [vp, sz]),
        "ocb_image_to_3d": (i32, [vp, sz, vp, vp]),
        "ocb_match_lists": (i32, [vp, sz, vp, sz, vp, vp, vp, sz, vp]),
        "ocb_match_lists_device": (i32, [vp, vp, vp, vp, vp, sz, vp, vp]),
        "ocb_score_models": (i32, [i32, vp, sz, vp, sz, dbl, vp, vp, vp, vp]),
        "ocb_residuals": (i32, [i32, vp, vp, sz, vp]),
        "ocb_fit_homography": (i32, [vp, sz, vp, sz, vp, vp]),
        "ocb_score_models_device": (i32, [i32, vp, sz, vp, vp, sz, dbl, vp, vp, vp, vp]),
        "ocb_prepare_correspondences_device": (i32, [vp, vp, sz, vp, vp, vp]),
    }
    for name, (res, args) in sigs.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    # diagnostics (include/ocb_probe.h)
    if hasattr(L, "ocb_probe_pipes"):
        L.ocb_probe_pipes.restype = i32
        L.ocb_probe_pipes.argtypes = [vp, i32]
    if hasattr(L, "ocb_probe_exact_math"):
        L.ocb_probe_exact_math.restype = i32
        L.ocb_probe_exact_math.argtypes = [u64, u64, C.c_uint32, vp]
    if hasattr(L, "ocb_probe_umma"):
        L.ocb_probe_umma.restype = i32
        L.ocb_probe_umma.argtypes = [i32, i32, i32, i32, i32, vp]
    _lib = L
    return L


def exported_symbols():
    """Names declared in include/ocb.h (used by the CPU test that checks the library exports all of them)."""
    import re
    hdr = open(os.path.join(os.path.dirname(PKG), "include", "ocb.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(ocb_[a-z0-9_]+)\s*\(", hdr)))


def check(rc):
    if rc != 0:
        raise OcbError(f"libocb error {rc}: {lib().ocb_last_error().decode()}")


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _rows(a):
    a = np.ascontiguousarray(a)
    if a.dtype != np.uint64:
        a = a.view(np.uint64)
    return a.reshape(-1, 8)


def init(device=0):
    check(lib().ocb_init(device))


def set_option(key, value):
    check(lib().ocb_set_option(key.encode(), int(value)))


def get_option(key):
    return int(lib().ocb_get_option(key.encode()))


def kernel_launches():
    return int(lib().ocb_kernel_launches())


# ---- K1, host buffers (the call a user of the C ABI makes) ----
def probe_exact_math(seed, n, exponent_spread):
    """include/ocb_probe.h: {divisions compared, differing, square roots compared, differing, divisions out of range}."""
    counts = np.zeros(5, dtype=np.uint64)
    check(lib().ocb_probe_exact_math(int(seed), int(n), int(exponent_spread), _ptr(counts)))
    return counts


def probe_umma(a_in_tmem, n, chains, rounds=512, ctas=148):
    """include/ocb_probe.h: cycles per tcgen05.mma.kind::i8 (M 128 x N n x K 32) over `chains` independent accumulators."""
    out = np.zeros(1, dtype=np.float64)
    check(lib().ocb_probe_umma(int(bool(a_in_tmem)), int(n), int(chains), int(rounds), int(ctas), _ptr(out)))
    return float(out[0])


def match_top2(q, c, cross_check=False, out=None, col_out=None):
    q, c = _rows(q), _rows(c)
    n1, n2 = len(q), len(c)
    if out is None:
        out = np.zeros(n1, TOP2_DTYPE)
    col = None
    if cross_check:
        col = col_out if col_out is not None else np.zeros(n2, np.uint32)
    check(lib().ocb_match_top2(_ptr(q), n1, _ptr(c), n2, _ptr(out), _ptr(col)))
    return (out, col) if cross_check else out


def match_top2_strided(base1, stride1, idx1, base2, stride2, idx2, cross_check=False):
    """Rows gathered from strided records: base*: uint8 arrays, row k = base[idx[k] * stride : +64] (idx None: k)."""
    n1 = len(idx1) if idx1 is not None else len(base1) // stride1
    n2 = len(idx2) if idx2 is not None else len(base2) // stride2
    i1 = None if idx1 is None else np.ascontiguousarray(idx1, dtype=np.uintp)
    i2 = None if idx2 is None else np.ascontiguousarray(idx2, dtype=np.uintp)
    out = np.zeros(n1, TOP2_DTYPE)
    col = np.zeros(n2, np.uint32) if cross_check else None
    check(lib().ocb_match_top2_strided(_ptr(base1), stride1, _ptr(i1), n1, _ptr(base2), stride2, _ptr(i2), n2,
                                       _ptr(out), _ptr(col)))
    return (out, col) if cross_check else out


# ---- K1, device buffers (torch tensors; inputs resident in HBM) ----
def match_top2_workspace_bytes(n1, n2, cross_check=False):
    return int(lib().ocb_match_top2_workspace_bytes(n1, n2, int(cross_check)))


def match_top2_device(d_q, n1, d_c, n2, d_out, d_col, d_ws, ws_bytes, stream):
    """All d_* are integer device addresses (tensor.data_ptr()); stream is a cudaStream_t address."""
    check(lib().ocb_match_top2_device(d_q, n1, d_c, n2, d_out, d_col, d_ws, ws_bytes, stream))


# ---- descriptor residency + batched pairs ----
def register_descriptors(set_id, rows):
    rows = _rows(rows)
    check(lib().ocb_register_descriptors(int(set_id), _ptr(rows), len(rows)))


def register_descriptors_batch(sets):
    """sets: [(set_id, base uint8 array, stride, idx or None, n)] -> one device allocation for all of them."""
    src = np.zeros(len(sets), SET_SOURCE_DTYPE)
    keep = []
    for i, (sid, base, stride, idx, n) in enumerate(sets):
        base = np.ascontiguousarray(base)
        idx = None if idx is None else np.ascontiguousarray(idx, dtype=np.uintp)
        keep += [base, idx]
        src[i] = (sid, base.ctypes.data if n else 0, stride, 0 if idx is None else idx.ctypes.data, n)
    check(lib().ocb_register_descriptors_batch(_ptr(src), len(src)))


def unregister_descriptors(set_id):
    check(lib().ocb_unregister_descriptors(int(set_id)))


def match_pairs(pairs, n_query_rows, out=None):
    """pairs: [(query_set, candidate_set)]; n_query_rows[p] = rows of pair p's query set. Returns (out, offsets)."""
    pa = np.zeros(len(pairs), PAIR_DTYPE)
    for i, (a, b) in enumerate(pairs):
        pa[i] = (a, b)
    offs = np.zeros(len(pairs), np.uint64)
    if len(pairs):
        offs[1:] = np.cumsum(np.asarray(n_query_rows[:-1], np.uint64))
    total = int(np.sum(np.asarray(n_query_rows, np.uint64)))
    if out is None:
        out = np.zeros(total, TOP2_DTYPE)
    check(lib().ocb_match_pairs(_ptr(pa), len(pa), _ptr(out), _ptr(offs)))
    return out, offs


MATCH_DTYPE = np.dtype([("query_k", np.uint32), ("best_k", np.uint32), ("best_d", np.uint32)])
CAMERA_DTYPE = np.dtype([("focal_length_pixels", "<f8"), ("principal_point", "<f8", 2), ("radial_distortion", "<f8", 3),
                         ("tangential_distortion", "<f8", 2), ("projection_planar", "<i4"), ("reserved", "<i4")])
IMAGE_SOURCE_DTYPE = np.dtype([("set_id", "<u8"), ("rows", "<u8"), ("stride", "<u8"), ("idx", "<u8"), ("n", "<u8"),
                               ("xy", "<u8"), ("xy_stride", "<u8"), ("camera", CAMERA_DTYPE)])
MATCH_SET_DTYPE = np.dtype([("set_1", "<u8"), ("set_2", "<u8"), ("matches", "<u8"), ("n", "<u8"), ("order", "<u8"),
                            ("corr_out", "<u8")])


def camera(cam8, planar=True):
    """cam8 = (f, ppx, ppy, k1, k2, k3, p1, p2) -> ocb_camera record."""
    c = np.zeros((), CAMERA_DTYPE)
    cam8 = np.asarray(cam8, np.float64)
    c["focal_length_pixels"] = cam8[0]
    c["principal_point"] = cam8[1:3]
    c["radial_distortion"] = cam8[3:6]
    c["tangential_distortion"] = cam8[6:8]
    c["projection_planar"] = 1 if planar else 0
    return c


def match_pairs_ratio(pairs, capacity):
    """ocb_match_pairs + device ratio test + compaction -> (matches [total] MATCH_DTYPE, offsets [n_pairs + 1])."""
    pa = np.zeros(len(pairs), PAIR_DTYPE)
    for i, (a, b) in enumerate(pairs):
        pa[i] = (a, b)
    out = np.zeros(max(int(capacity), 1), MATCH_DTYPE)
    offs = np.zeros(len(pairs) + 1, np.uint64)
    check(lib().ocb_match_pairs_ratio(_ptr(pa), len(pa), _ptr(out), int(capacity), _ptr(offs)))
    return out[:int(offs[-1])], offs


def match_pairs_sorted(pairs, capacity, want_quality_order=True):
    """ocb_match_pairs_ratio + the reference's std::sort on the device -> (matches in the reference's final order,
    offsets [n_pairs + 1], quality_order or None)."""
    pa = np.zeros(len(pairs), PAIR_DTYPE)
    for i, (a, b) in enumerate(pairs):
        pa[i] = (a, b)
    out = np.zeros(max(int(capacity), 1), MATCH_DTYPE)
    offs = np.zeros(len(pairs) + 1, np.uint64)
    qo = np.zeros(max(int(capacity), 1), np.uint32) if want_quality_order else None
    check(lib().ocb_match_pairs_sorted(_ptr(pa), len(pa), _ptr(out), int(capacity), _ptr(offs), _ptr(qo)))
    total = int(offs[-1])
    return out[:total], offs, (qo[:total] if want_quality_order else None)


def register_images_batch(images):
    """images: [(set_id, rows [n][8] u64, xy [n][2] f64 or None, cam8 or None)]: rows, keypoints and camera model."""
    src = np.zeros(len(images), IMAGE_SOURCE_DTYPE)
    keep = []
    for i, (sid, rows, xy, cam8) in enumerate(images):
        rows = _rows(rows)
        xy = None if xy is None else np.ascontiguousarray(xy, np.float64).reshape(-1, 2)
        keep += [rows, xy]
        src[i]["set_id"], src[i]["n"] = sid, len(rows)
        src[i]["rows"], src[i]["stride"] = (rows.ctypes.data if len(rows) else 0), 64
        if xy is not None:
            assert len(xy) == len(rows)
            src[i]["xy"], src[i]["xy_stride"] = (xy.ctypes.data if len(xy) else 0), 16
        if cam8 is not None:
            src[i]["camera"] = camera(cam8)
    check(lib().ocb_register_images_batch(_ptr(src), len(src)))


def corr_bind_batch_matches(sets):
    """sets: [(set_1, set_2, matches MATCH_DTYPE [n], order uint32 [n] or None)] -> list of [n][7] correspondence rows
    computed on the device (K6); the sets stay bound on this thread for ocb_score_requests."""
    ms = np.zeros(len(sets), MATCH_SET_DTYPE)
    keep, outs = [], []
    for i, (s1, s2, matches, order) in enumerate(sets):
        matches = np.ascontiguousarray(matches, MATCH_DTYPE)
        order = None if order is None else np.ascontiguousarray(order, np.uint32)
        out = np.full((len(matches), 7), np.nan)
        keep += [matches, order]
        outs.append(out)
        ms[i] = (s1, s2, matches.ctypes.data if len(matches) else 0, len(matches),
                 0 if order is None else order.ctypes.data, out.ctypes.data if len(matches) else 0)
    check(lib().ocb_corr_bind_batch_matches(_ptr(ms), len(ms)))
    return outs


def image_to_3d(xy, cam8, planar=True):
    xy = np.ascontiguousarray(xy, np.float64).reshape(-1, 2)
    rays = np.zeros((len(xy), 3))
    cam = camera(cam8, planar)
    check(lib().ocb_image_to_3d(_ptr(xy), len(xy), cam.ctypes.data_as(C.c_void_p), _ptr(rays)))
    return rays


# ---- K4: candidate lists ----
def match_lists(q, c, list_query, list_begin, list_candidates):
    """Top-2 over per-query candidate lists (CSR). -> TOP2 records, best_k = position within the list."""
    q, c = _rows(q), _rows(c)
    list_query = np.ascontiguousarray(list_query, np.uint32)
    list_begin = np.ascontiguousarray(list_begin, np.uint64)
    list_candidates = np.ascontiguousarray(list_candidates, np.uint32)
    nl = len(list_query)
    assert len(list_begin) == nl + 1
    out = np.zeros(nl, TOP2_DTYPE)
    check(lib().ocb_match_lists(_ptr(q), len(q), _ptr(c), len(c), _ptr(list_query), _ptr(list_begin),
                                _ptr(list_candidates), nl, _ptr(out)))
    return out


def match_lists_device(d_q, d_c, d_list_query, d_list_begin, d_list_candidates, n_lists, d_out, stream):
    check(lib().ocb_match_lists_device(d_q, d_c, d_list_query, d_list_begin, d_list_candidates, n_lists, d_out, stream))


# ---- K2 / K3 ----
def score_models(kind, models, corr, thr, order=None, want_bits=True):
    models = np.ascontiguousarray(models, np.float64).reshape(-1, 18)
    corr = np.ascontiguousarray(corr, np.float64).reshape(-1, 7)
    h, n = len(models), len(corr)
    score, count = np.zeros(h, np.float64), np.zeros(h, np.uint32)
    bits = np.zeros((h, (n + 31) // 32), np.uint32) if want_bits else None
    order = None if order is None else np.ascontiguousarray(order, np.uint32)
    check(lib().ocb_score_models(kind, _ptr(models), h, _ptr(corr), n, float(thr), _ptr(order), _ptr(score),
                                 _ptr(count), _ptr(bits)))
    return score, count, bits


def residuals(kind, model18, corr):
    model18 = np.ascontiguousarray(model18, np.float64).reshape(18)
    corr = np.ascontiguousarray(corr, np.float64).reshape(-1, 7)
    e = np.zeros(len(corr), np.float64)
    check(lib().ocb_residuals(kind, _ptr(model18), _ptr(corr), len(corr), _ptr(e)))
    return e


def fit_homography(corr, samples):
    """Minimal-sample homography fits on the device -> (models [h][18], degenerate [h] bool)."""
    corr = np.ascontiguousarray(corr, np.float64).reshape(-1, 7)
    samples = np.ascontiguousarray(samples, np.uint32).reshape(-1, 4)
    h = len(samples)
    models, deg = np.zeros((h, 18), np.float64), np.zeros(h, np.uint8)
    check(lib().ocb_fit_homography(_ptr(corr), len(corr), _ptr(samples), h, _ptr(models), _ptr(deg)))
    return models, deg.astype(bool)


# ---- K2 / K3, device buffers ----
def prepare_correspondences_device(d_corr7, d_order, n, d_corr4, d_pos, stream):
    check(lib().ocb_prepare_correspondences_device(d_corr7, d_order, n, d_corr4, d_pos, stream))


def score_models_device(kind, d_models, h, d_corr4, d_pos, n, thr, d_score, d_count, d_bits, stream):
    check(lib().ocb_score_models_device(kind, d_models, h, d_corr4, d_pos, n, float(thr), d_score, d_count, d_bits,
                                        stream))
