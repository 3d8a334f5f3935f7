"""TEST INFRASTRUCTURE ONLY. ctypes binding of oracle/_ref/liboc_ref_io.so: the reference's OWN graph.json reader and
writer (src/io/{serialize,deserialize}_MeasurementGraph.cpp, src/io/base64.c) compiled in place by oracle/Makefile
(target ref_io), plus numpy restatements of the two bit-packing helpers. Only tests/ and tests/golden/ scripts import
this; the product never does."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_ref", "liboc_ref_io.so")
_u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_lib = None


def build():
    """Only possible where /root/reference and a rapidjson header tree exist (the build container)."""
    if os.path.isdir("/root/reference"):
        subprocess.run(["make", "-C", HERE, "ref_io"], check=False, stdout=subprocess.DEVNULL,
                       stderr=subprocess.DEVNULL)
    return os.path.exists(SO)


def available():
    return os.path.exists(SO)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(SO)
        sz, vp = C.c_size_t, C.c_void_p
        L.oc_ref_io_roundtrip.argtypes = [C.c_char_p, sz, vp, sz, C.POINTER(C.c_int)]
        L.oc_ref_io_roundtrip.restype = sz
        L.oc_ref_graph_new.restype = vp
        L.oc_ref_graph_free.argtypes = [vp]
        L.oc_ref_graph_free.restype = None
        L.oc_ref_graph_add_node.argtypes = [vp, C.c_char_p, _f64p, C.c_int64, _f64p, _u64p, C.POINTER(C.c_char_p),
                                            _f64p, _f64p, _f32p, _u64p, sz, sz]
        L.oc_ref_graph_add_node.restype = C.c_uint64
        L.oc_ref_graph_add_edge.argtypes = [vp, C.c_uint64, C.c_uint64, _u64p, _u64p, _f64p, sz, _f64p, _u64p, sz,
                                            C.c_int, _f64p, _f64p]
        L.oc_ref_graph_add_edge.restype = C.c_uint64
        L.oc_ref_graph_serialize.argtypes = [vp, vp, sz]
        L.oc_ref_graph_serialize.restype = sz
        L.Base64encode.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
        L.Base64encode_len.argtypes = [C.c_int]
        L.Base64decode.argtypes = [C.c_char_p, C.c_char_p]
        L.Base64decode_len.argtypes = [C.c_char_p]
        _lib = L
    return _lib


def roundtrip(text):
    """reference deserialize() then serialize(); None when deserialize() returns false.
    -> (rewritten bytes, graph-equal-after-another-round-trip)"""
    text = text.encode() if isinstance(text, str) else bytes(text)
    eq = C.c_int(0)
    n = lib().oc_ref_io_roundtrip(text, len(text), None, 0, None)
    if n == 0:
        return None
    buf = C.create_string_buffer(n + 1)
    lib().oc_ref_io_roundtrip(text, len(text), buf, n, C.byref(eq))
    return buf.raw[:n], bool(eq.value)


def base64_encode(data):
    """Base64encode of src/io/base64.c"""
    data = bytes(data)
    out = C.create_string_buffer(lib().Base64encode_len(len(data)) + 1)
    n = lib().Base64encode(out, data, len(data))
    return out.raw[:n - 1]


def base64_decode(text):
    """Base64decode of src/io/base64.c (text must not hold NUL)"""
    text = bytes(text)
    out = C.create_string_buffer(lib().Base64decode_len(text) + 4)
    n = lib().Base64decode(out, text)
    return out.raw[:n]


class RefGraph:
    """A MeasurementGraph of the reference, filled through its own addNode/addEdge."""
    STR8 = ("make", "model", "serial_no", "lens_make", "lens_model", "datum", "timestamp", "datestamp")

    def __init__(self):
        self.h = lib().oc_ref_graph_new()

    def add_node(self, path, pose7, model_id, camera8, dims2, xy, strength, rows, num_sparse, strings=None,
                 capture9=None):
        strings = strings or {}
        arr = (C.c_char_p * 8)(*[strings.get(k, "").encode() for k in self.STR8])
        cap = np.full(9, np.nan) if capture9 is None else np.ascontiguousarray(capture9, np.float64)
        xy = np.ascontiguousarray(xy, np.float64).reshape(-1, 2)
        return lib().oc_ref_graph_add_node(self.h, path.encode(), np.ascontiguousarray(pose7, np.float64), model_id,
                                           np.ascontiguousarray(camera8, np.float64),
                                           np.ascontiguousarray(dims2, np.uint64), arr, cap, xy,
                                           np.ascontiguousarray(strength, np.float32),
                                           np.ascontiguousarray(rows).view(np.uint64).reshape(-1, 8), len(xy),
                                           int(num_sparse))

    def add_edge(self, source, dest, matches, inlier_pixels, inlier_idx, relation_type, relation9, poses32):
        i1, i2, d = (np.ascontiguousarray(matches[0], np.uint64), np.ascontiguousarray(matches[1], np.uint64),
                     np.ascontiguousarray(matches[2], np.float64))
        px = np.ascontiguousarray(inlier_pixels, np.float64).reshape(-1, 4)
        ix = np.ascontiguousarray(inlier_idx, np.uint64).reshape(-1, 3)
        return lib().oc_ref_graph_add_edge(self.h, source, dest, i1, i2, d, len(d), px, ix, len(px),
                                           int(relation_type), np.ascontiguousarray(relation9, np.float64).reshape(9),
                                           np.ascontiguousarray(poses32, np.float64).reshape(32))

    def serialize(self):
        n = lib().oc_ref_graph_serialize(self.h, None, 0)
        buf = C.create_string_buffer(n + 1)
        lib().oc_ref_graph_serialize(self.h, buf, n)
        return buf.raw[:n]

    def __del__(self):
        if getattr(self, "h", None):
            lib().oc_ref_graph_free(self.h)
            self.h = None


# ---- numpy restatements (the oracle "port" of the two helpers) ----
def bitset_to_bytes(rows):
    """serialize_MeasurementGraph.cpp:20-27 for rows [n][8] u64 (memory image of std::bitset<486>) -> [n][61] u8:
    result[j >> 3] |= bit j << (j & 7), j < 486."""
    rows = np.ascontiguousarray(rows).view(np.uint64).reshape(-1, 8)
    bits = ((rows[:, :, None] >> np.arange(64, dtype=np.uint64)[None, None, :]) & np.uint64(1)).astype(np.uint8)
    bits = bits.reshape(len(rows), 512)[:, :486]
    bits = np.concatenate([bits, np.zeros((len(rows), 2), np.uint8)], axis=1)
    return np.packbits(bits, axis=1, bitorder="little")


def bitset_from_bytes(b61):
    """deserialize_MeasurementGraph.cpp:17-24: [n][61] u8 -> rows [n][8] u64, bits 486.. zero"""
    b61 = np.ascontiguousarray(b61, np.uint8).reshape(-1, 61)
    bits = np.unpackbits(b61, axis=1, bitorder="little")[:, :486]
    bits = np.concatenate([bits, np.zeros((len(b61), 26), np.uint8)], axis=1).reshape(len(b61), 8, 64)
    return (bits.astype(np.uint64) << np.arange(64, dtype=np.uint64)[None, None, :]).sum(axis=2, dtype=np.uint64)
