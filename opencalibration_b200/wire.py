"""ctypes binding of include/ocb_wire.h (libocb_host.so): the reference's graph.json wire format for the matching path
-- descriptors as 61 little-endian-bit-packed bytes in base64, matches as [index_1, index_2, distance] -- read into the
device layout and written back byte-compatibly (src/io/{serialize,deserialize}_MeasurementGraph.cpp)."""
import ctypes as C

import numpy as np

from .capi import OcbError
from .host import lib as _host_lib

_ready = False
_u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")


def lib():
    global _ready
    L = _host_lib()
    if _ready:
        return L
    sz, vp, i32, cp = C.c_size_t, C.c_void_p, C.c_int, C.c_char_p
    L.ocbw_last_error.restype = cp
    L.ocbw_descriptor_encode.argtypes = [_u64p, vp]
    L.ocbw_descriptor_decode.argtypes = [cp, sz, _u64p]
    L.ocbw_base64_encode.argtypes = [cp, sz, vp, sz]
    L.ocbw_base64_encode.restype = sz
    L.ocbw_base64_decode.argtypes = [cp, sz, vp, sz]
    L.ocbw_base64_decode.restype = sz
    L.ocbw_format_double.argtypes = [C.c_double, vp]
    L.ocbw_format_double.restype = sz
    L.ocbw_parse_double.argtypes = [cp, sz, C.POINTER(C.c_double)]
    L.ocbw_graph_parse.argtypes = [cp, sz]
    L.ocbw_graph_parse.restype = vp
    L.ocbw_graph_create.restype = vp
    L.ocbw_graph_free.argtypes = [vp]
    L.ocbw_graph_free.restype = None
    L.ocbw_graph_num_nodes.argtypes = [vp]
    L.ocbw_graph_num_nodes.restype = sz
    L.ocbw_graph_num_edges.argtypes = [vp]
    L.ocbw_graph_num_edges.restype = sz
    L.ocbw_graph_node_info.argtypes = [vp, sz, vp, vp, vp, vp, vp, vp]
    L.ocbw_graph_node_features.argtypes = [vp, sz, vp, vp, vp]
    L.ocbw_graph_add_node.argtypes = [vp, vp, i32, cp, _f64p, _u64p, vp, _f64p, _f32p, _u64p, sz, sz]
    L.ocbw_graph_edge_info.argtypes = [vp, sz, vp, vp, vp, vp, vp, vp, vp, vp]
    L.ocbw_graph_edge_matches.argtypes = [vp, sz, vp, vp, vp, vp, vp]
    L.ocbw_graph_add_edge.argtypes = [vp, C.c_uint64, C.c_uint64, _u64p, _u64p, _f64p, sz, _f64p, _u64p, sz, i32,
                                      _f64p, _f64p, vp]
    L.ocbw_graph_serialize.argtypes = [vp, vp, sz]
    L.ocbw_graph_serialize.restype = sz
    L.ocbw_graph_link.argtypes = [vp, _u64p, sz, i32, i32, vp]
    _ready = True
    return L


def _err():
    return OcbError("wire: " + lib().ocbw_last_error().decode())


def _check(rc):
    if rc != 0:
        raise _err()


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def format_double(v):
    buf = C.create_string_buffer(40)
    n = lib().ocbw_format_double(float(v), buf)
    return buf.raw[:n].decode()


def parse_double(text):
    b = text.encode() if isinstance(text, str) else text
    v = C.c_double()
    _check(lib().ocbw_parse_double(b, len(b), C.byref(v)))
    return v.value


def base64_encode(data):
    data = bytes(data)
    out = C.create_string_buffer((len(data) + 2) // 3 * 4 + 4)
    n = lib().ocbw_base64_encode(data, len(data), out, len(out))
    return out.raw[:n]


def base64_decode(text):
    text = text.encode() if isinstance(text, str) else bytes(text)
    out = C.create_string_buffer((len(text) + 3) // 4 * 3 + 4)
    n = lib().ocbw_base64_decode(text, len(text), out, len(out))
    return out.raw[:n]


def descriptor_encode(row):
    row = np.ascontiguousarray(row).view(np.uint64).reshape(8)
    out = C.create_string_buffer(84)
    _check(lib().ocbw_descriptor_encode(row, out))
    return out.raw[:84]


def descriptor_decode(text):
    text = text.encode() if isinstance(text, str) else bytes(text)
    row = np.zeros(8, np.uint64)
    _check(lib().ocbw_descriptor_decode(text, len(text), row))
    return row


class Graph:
    """A graph.json document (image nodes with features + camera model, edges = camera_relations)."""

    def __init__(self, json_text=None):
        L = lib()
        if json_text is None:
            self.h = L.ocbw_graph_create()
        else:
            b = json_text.encode() if isinstance(json_text, str) else bytes(json_text)
            self.h = L.ocbw_graph_parse(b, len(b))
            if not self.h:
                raise _err()

    def close(self):
        if getattr(self, "h", None):
            lib().ocbw_graph_free(self.h)
            self.h = None

    __del__ = close

    @property
    def num_nodes(self):
        return lib().ocbw_graph_num_nodes(self.h)

    @property
    def num_edges(self):
        return lib().ocbw_graph_num_edges(self.h)

    def node(self, i, features=True):
        nid, nf, ns = C.c_uint64(), C.c_size_t(), C.c_size_t()
        cam, dims, pose = np.zeros(8), np.zeros(2, np.uint64), np.zeros(7)
        _check(lib().ocbw_graph_node_info(self.h, i, C.byref(nid), C.byref(nf), C.byref(ns), _p(cam), _p(dims),
                                          _p(pose)))
        out = dict(id=nid.value, n_features=nf.value, num_sparse_features=ns.value, camera=cam, dims=dims, pose=pose)
        if features:
            n = nf.value
            xy, st, rows = np.zeros((n, 2)), np.zeros(n, np.float32), np.zeros((n, 8), np.uint64)
            _check(lib().ocbw_graph_node_features(self.h, i, _p(xy), _p(st), _p(rows)))
            out.update(xy=xy, strength=st, rows=rows)
        return out

    def add_node(self, camera, dims, xy, strength, rows, num_sparse=None, node_id=None, path="", pose=None):
        """node_id None: the id is drawn like MeasurementGraph::addNode draws it."""
        xy = np.ascontiguousarray(xy, np.float64).reshape(-1, 2)
        st = np.ascontiguousarray(strength, np.float32)
        rows = np.ascontiguousarray(rows).view(np.uint64).reshape(-1, 8)
        nid = C.c_uint64(0 if node_id is None else node_id)
        pose = None if pose is None else np.ascontiguousarray(pose, np.float64)
        _check(lib().ocbw_graph_add_node(self.h, C.byref(nid), int(node_id is None), path.encode(),
                                         np.ascontiguousarray(camera, np.float64),
                                         np.ascontiguousarray(dims, np.uint64), _p(pose), xy, st, rows, len(xy),
                                         len(xy) if num_sparse is None else int(num_sparse)))
        return nid.value

    def edge(self, i):
        eid, src, dst, nm, ni, rt = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_size_t(), C.c_size_t(), C.c_int()
        rel, poses = np.zeros(9), np.zeros(32)
        _check(lib().ocbw_graph_edge_info(self.h, i, C.byref(eid), C.byref(src), C.byref(dst), C.byref(nm), C.byref(ni),
                                          C.byref(rt), _p(rel), _p(poses)))
        i1, i2, d = np.zeros(nm.value, np.uint64), np.zeros(nm.value, np.uint64), np.zeros(nm.value)
        px, ix = np.zeros((ni.value, 4)), np.zeros((ni.value, 3), np.uint64)
        _check(lib().ocbw_graph_edge_matches(self.h, i, _p(i1), _p(i2), _p(d), _p(px), _p(ix)))
        return dict(id=eid.value, source=src.value, dest=dst.value, relation_type=rt.value, relation=rel.reshape(3, 3),
                    poses=poses.reshape(4, 8), matches=(i1, i2, d), inlier_pixels=px, inlier_idx=ix)

    def add_edge(self, source, dest, matches, inlier_pixels, inlier_idx, relation_type, relation, poses):
        i1, i2, d = (np.ascontiguousarray(matches[0], np.uint64), np.ascontiguousarray(matches[1], np.uint64),
                     np.ascontiguousarray(matches[2], np.float64))
        px = np.ascontiguousarray(inlier_pixels, np.float64).reshape(-1, 4)
        ix = np.ascontiguousarray(inlier_idx, np.uint64).reshape(-1, 3)
        eid = C.c_uint64()
        _check(lib().ocbw_graph_add_edge(self.h, source, dest, i1, i2, d, len(d), px, ix, len(px), int(relation_type),
                                         np.ascontiguousarray(relation, np.float64).reshape(9),
                                         np.ascontiguousarray(poses, np.float64).reshape(32), C.byref(eid)))
        return eid.value

    def serialize(self):
        n = lib().ocbw_graph_serialize(self.h, None, 0)
        buf = C.create_string_buffer(n + 1)
        lib().ocbw_graph_serialize(self.h, buf, n)
        return buf.raw[:n]

    def link(self, pairs, threads=0, run_ransac=True):
        """LinkStage over the document's own features on the GPU; results become edges (pair order)."""
        pr = np.ascontiguousarray(pairs, np.uint64).reshape(-1, 2)
        sec = np.zeros(4)
        _check(lib().ocbw_graph_link(self.h, pr, len(pr), int(threads), int(run_ransac), _p(sec)))
        return dict(seconds_subsample_upload=sec[0], seconds_match_gpu=sec[1], seconds_tail=sec[2],
                    seconds_total=sec[3])
