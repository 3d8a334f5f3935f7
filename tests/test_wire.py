"""Wire format of the matching path (host/graph_wire.hpp, include/ocb_wire.h): descriptors as 61
LE-bit-packed bytes in base64, matches as [index_1, index_2, distance], inside the reference's graph.json.
Pinned to (a) golden vectors written by the reference's OWN serializer object code (tests/golden/wire_vectors.npz,
wire_graph.json; generator make_wire_vectors.py) and (b), where oracle/_ref/liboc_ref_io.so is present, to that object
code directly: documents built through the reference's addNode/addEdge + serialize() must come out byte-identical
from the product, and product output must survive the reference's deserialize() -> serialize() unchanged
(the property test/test_serialize_deserialize.cpp:24-64 checks)."""
import base64
import os
import re

import numpy as np
import pytest

import oc_ref_io as R
from opencalibration_b200 import capi, wire

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def vectors(hostlib):
    z = np.load(os.path.join(GOLDEN, "wire_vectors.npz"))

    def items(name):
        b, o = z[name + "_bytes"].tobytes(), z[name + "_offsets"]
        return [b[o[i]:o[i + 1]] for i in range(len(o) - 1)]
    return dict(bits=z["double_bits"], text=items("double_text"), plain=items("b64_plain"), coded=items("b64_coded"),
                odd=items("b64_odd"), odd_plain=items("b64_odd_plain"))


@pytest.fixture(scope="module")
def golden_graph(hostlib):
    return open(os.path.join(GOLDEN, "wire_graph.json"), "rb").read()


@pytest.fixture(scope="module")
def ref_io():
    if not R.available() and not R.build():
        pytest.skip("oracle/_ref/liboc_ref_io.so not built and /root/reference absent")
    return R


def test_header_symbols_exported(hostlib):
    hdr = open(os.path.join(os.path.dirname(GOLDEN), "..", "include", "ocb_wire.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = sorted(set(re.findall(r"\b(ocbw_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 18
    L = wire.lib()
    for n in names:
        assert hasattr(L, n), f"libocb_host.so does not export {n}"


def test_format_double_equals_reference_writer(vectors):
    d = vectors["bits"].view(np.float64)
    assert len(d) == len(vectors["text"]) > 7000
    bad = [(float(v), t) for v, t in zip(d, vectors["text"]) if wire.format_double(v).encode() != t]
    assert not bad, bad[:5]


def test_parse_double_round_trips_every_golden_text(vectors):
    for bits, t in zip(vectors["bits"], vectors["text"]):
        got = np.array([wire.parse_double(t)]).view(np.uint64)[0]
        want = bits if not np.isnan(np.array([bits]).view(np.float64)[0]) else got  # NaN payloads are not kept
        assert got == want, t
    for t in ("1", "-0", "12e3", "1E-2", "0.5", "-Infinity", "Inf", "NaN", "18446744073709551616"):
        wire.parse_double(t)
    for t in ("", "-", "01", "1.", ".5", "1e", "+1", "0x10", "1,0", "nan", "-nan", "1e999", "1e+", "--1", "1.5.2"):
        with pytest.raises(capi.OcbError):
            wire.parse_double(t)


def test_base64_equals_reference(vectors):
    for p, c in zip(vectors["plain"], vectors["coded"]):
        assert wire.base64_encode(p) == c == base64.b64encode(p)
        assert wire.base64_decode(c) == p
    for t, p in zip(vectors["odd"], vectors["odd_plain"]):
        assert wire.base64_decode(t) == p, t


def test_descriptor_packing(hostlib):
    rng = np.random.default_rng(0)
    rows = rng.integers(0, 2 ** 64, (300, 8), dtype=np.uint64)
    rows[:, 7] &= np.uint64((1 << 38) - 1)  # bits 486.. are zero in a std::bitset<486>
    rows[0] = 0
    rows[1] = np.uint64(2 ** 64 - 1)
    rows[1, 7] = np.uint64((1 << 38) - 1)
    packed = R.bitset_to_bytes(rows)  # numpy restatement of serialize :20-27
    assert packed.shape == (300, 61) and np.array_equal(R.bitset_from_bytes(packed), rows)
    for r, b in zip(rows, packed):
        text = wire.descriptor_encode(r)
        assert len(text) == 84 and text == base64.b64encode(b.tobytes())
        assert np.array_equal(wire.descriptor_decode(text), r)
    dirty = rows[5].copy()
    dirty[7] |= np.uint64(1 << 50)  # stray bits above 486 never reach the wire
    assert wire.descriptor_encode(dirty) == wire.descriptor_encode(rows[5])
    for bad in (b"", b"AAAA", wire.descriptor_encode(rows[2])[:80], b"A" + wire.descriptor_encode(rows[2])):
        with pytest.raises(capi.OcbError):
            wire.descriptor_decode(bad)
    # like Base64decode, decoding stops at the padding: text after it is ignored
    assert np.array_equal(wire.descriptor_decode(wire.descriptor_encode(rows[2]) + b"AAAA"), rows[2])


def test_golden_graph_reads_and_writes_back_identically(golden_graph):
    g = wire.Graph(golden_graph)
    assert g.num_nodes == 3 and g.num_edges == 3
    assert g.serialize() == golden_graph
    # what the reader hands to the device side
    n0 = g.node(0)
    assert n0["n_features"] == 40 and n0["num_sparse_features"] == 35 and n0["rows"].shape == (40, 8)
    assert (n0["rows"][:, 7] >> np.uint64(38)).max() == 0
    assert n0["camera"][0] == 3000.0 and tuple(n0["dims"]) == (4000, 3000)
    e = [g.edge(i) for i in range(3)]
    by_sizes = sorted((len(x["matches"][2]), len(x["inlier_idx"])) for x in e)
    assert by_sizes == [(5, 5), (9, 0), (12, 6)]
    for x in e:
        d = x["matches"][2]
        assert np.all(np.diff(d) <= 0) and np.array_equal(np.rint(d * 486) / 486, d)  # sorted, multiples of 1/486
        if len(x["inlier_idx"]) == 0:
            assert x["relation_type"] == 2 and np.isnan(x["poses"][:, 1:]).all()


def test_camera_models_are_shared_by_model_id(golden_graph):
    """deserialize :84-110: a node whose model id appeared earlier in the document gets that earlier model."""
    text = golden_graph.decode()
    assert text.count('"focal_length": 3000.0') == 2 and text.count('"focal_length": 3030.0') == 1
    # give the LAST node that carries model id 0 another focal length: the reader must ignore it
    head, sep, tail = text.rpartition('"focal_length": 3000.0')
    g = wire.Graph(head + '"focal_length": 1234.5' + tail)
    assert sorted(g.node(i, False)["camera"][0] for i in range(3)) == [3000.0, 3000.0, 3030.0]
    assert g.serialize() == golden_graph


def test_reader_is_order_free_and_tolerant_like_a_dom_lookup(golden_graph):
    import json
    doc = json.loads(golden_graph.decode().replace("NaN", '"@nan"'))
    # reversed member order everywhere + an unknown member: the reference looks members up by name
    def rev(o, container=False):
        if isinstance(o, dict):
            out = {k: rev(v, k in ("nodes", "edges") and not container) for k, v in reversed(list(o.items()))}
            if not container:  # "nodes" / "edges" map ids to objects: every member there is a node / an edge
                out["unknown_member"] = [1, {"a": None}, True]
            return out
        if isinstance(o, list):
            return [rev(v) for v in o]
        return o
    text = json.dumps(rev(doc)).replace('"@nan"', "NaN")
    g = wire.Graph(text)
    # same graph (node order in the text does not matter: written sorted by id) except the model-sharing winner
    a, b = wire.Graph(golden_graph), g
    assert a.num_nodes == b.num_nodes and a.num_edges == b.num_edges
    ids_a = sorted(a.node(i, False)["id"] for i in range(3))
    ids_b = sorted(b.node(i, False)["id"] for i in range(3))
    assert ids_a == ids_b
    ea = {a.edge(i)["id"]: a.edge(i) for i in range(3)}
    for i in range(3):
        x = b.edge(i)
        y = ea[x["id"]]
        assert all(np.array_equal(p, q) for p, q in zip(x["matches"], y["matches"]))
        assert np.array_equal(x["relation"], y["relation"]) and np.array_equal(x["poses"], y["poses"], equal_nan=True)


@pytest.mark.parametrize("text,why", [
    ("", "empty"), ("[]", "not an object"), ('{"version": 2, "nodes": {}, "edges": {}}', "version"),
    ('{"nodes": {}, "edges": {}}', "no version"), ('{"version": 1, "nodes": {}}', "no edges"),
    ('{"version": 1, "nodes": {}, "edges": {}} x', "trailing"), ('{"version": 1, "nodes": {"1": {}}, "edges": {}}',
                                                                  "node members"),
    ('{"version": 1, "nodes": {}, "edges": {"7": {"source": "1"}}}', "edge members"),
    ('{"version": 1.0, "nodes": {}, "edges": {}}', "version is not an integer (IsInt64)"),
])
def test_reader_rejects(hostlib, text, why):
    with pytest.raises(capi.OcbError):
        wire.Graph(text)


def test_truncated_and_corrupted_documents_fail_cleanly(golden_graph):
    rng = np.random.default_rng(1)
    for cut in rng.integers(1, len(golden_graph) - 1, 60):
        with pytest.raises(capi.OcbError):
            wire.Graph(golden_graph[:cut])
    bad = golden_graph.replace(b'"descriptor": "', b'"descriptor": "AAAA', 1)
    with pytest.raises(capi.OcbError, match="61"):
        wire.Graph(bad)


def test_num_sparse_features_beyond_the_features_is_rejected(hostlib, golden_graph):
    # the reference trusts this count (link_stage.cpp:63-65); here a crafted file must not reach the matcher
    g = wire.Graph(golden_graph)
    n = g.node(0)["n_features"]
    g.close()
    text = golden_graph.decode()
    m = re.search(r'"num_sparse_features": (\d+)', text)
    assert m
    bad = text[:m.start(1)] + str(n + 1000000) + text[m.end(1):]
    with pytest.raises(capi.OcbError, match="num_sparse_features"):
        wire.Graph(bad.encode())
    g = wire.Graph()
    with pytest.raises(capi.OcbError, match="num_sparse_features"):
        g.add_node(np.zeros(8), np.array([10, 10], np.uint64), np.zeros((3, 2)), np.zeros(3, np.float32),
                   np.zeros((3, 8), np.uint64), num_sparse=4)
    g.close()


def test_subsample_count_beyond_the_features_means_all(hostlib, oracle):
    rng = np.random.default_rng(4)
    xy, st = rng.uniform(0, 500, (300, 2)), rng.uniform(0, 1, 300).astype(np.float32)
    assert np.array_equal(hostlib.spatially_subsample_feature_indices(xy, st, 20.0, 100000),
                          hostlib.spatially_subsample_feature_indices(xy, st, 20.0, 0))


def _random_graph(rng, n_nodes, n_edges, builders):
    """Same addNode / addEdge sequence into every builder (the reference's MeasurementGraph and the product's)."""
    cam = np.array([3000.0, 2000.5, 1500.25, -0.1, 0.01, 1e-4, 1e-5, -2e-5])
    dims = np.array([4000, 3000], np.uint64)
    ids = [[] for _ in builders]
    for k in range(n_nodes):
        n = int(rng.integers(0, 60))
        rows = rng.integers(0, 2 ** 64, (n, 8), dtype=np.uint64)
        rows[:, 7] &= np.uint64((1 << 38) - 1)
        xy = rng.uniform(-10, 6000, (n, 2))
        st = rng.uniform(0, 0.05, n).astype(np.float32)
        pose = rng.normal(size=7)
        for b, out in zip(builders, ids):
            out.append(b(k, pose, cam, dims, xy, st, rows, n // 2))
    for _ in range(n_edges):
        a, c = (int(v) for v in rng.integers(0, n_nodes, 2))
        nm = int(rng.integers(0, 30))
        ni = int(rng.integers(0, nm + 1))
        i1, i2 = rng.integers(0, 60, nm).astype(np.uint64), rng.integers(0, 60, nm).astype(np.uint64)
        dist = rng.integers(0, 487, nm) / 486
        px = rng.uniform(0, 4000, (ni, 4))
        ix = rng.integers(0, 60, (ni, 3)).astype(np.uint64)
        H = rng.normal(size=(3, 3)) * 10.0 ** rng.integers(-8, 8)
        poses = rng.normal(size=(4, 8))
        poses[:, 0] = rng.integers(-3, 200, 4)
        if rng.random() < 0.3:
            poses[:, 1:] = np.nan
        rt = int(rng.integers(0, 3))
        for b, out in zip(builders, ids):
            b.__self__.add_edge(out[a], out[c], (i1, i2, dist), px, ix, rt, H, poses)
    return ids


def test_same_calls_give_the_same_bytes_as_the_reference(ref_io, hostlib):
    rng = np.random.default_rng(7)
    for trial in range(4):
        rg, pg = ref_io.RefGraph(), wire.Graph()

        def ref_add(k, pose, cam, dims, xy, st, rows, ns, rg=rg):
            return rg.add_node("img%d" % k, pose, 0, cam, dims, xy, st, rows, ns)  # one shared camera model (id 0)
        ref_add.__self__ = rg

        def prod_add(k, pose, cam, dims, xy, st, rows, ns, pg=pg):
            return pg.add_node(cam, dims, xy, st, rows, num_sparse=ns, path="img%d" % k, pose=pose)
        prod_add.__self__ = pg
        ids = _random_graph(rng, 2 + trial * 3, trial * 7, [ref_add, prod_add])
        assert ids[0] == ids[1]  # node ids are draws of the same default-seeded generator (graph.hpp:73-84)
        want, got = rg.serialize(), pg.serialize()
        assert got == want
        # and through the readers: product reads reference text, reference reads product text
        assert wire.Graph(want).serialize() == want
        again, equal = ref_io.roundtrip(got)
        assert again == got and equal


def test_reference_reader_accepts_the_golden_fixture_unchanged(ref_io, golden_graph):
    again, equal = ref_io.roundtrip(golden_graph)
    assert again == golden_graph and equal and wire.Graph(golden_graph).serialize() == golden_graph
    assert wire.Graph(again).serialize() == again


def test_added_edges_continue_the_reference_id_sequence(ref_io, golden_graph):
    """addEdge on a freshly deserialized graph: fresh generator, ids that collide with existing ones are skipped."""
    rg = ref_io.RefGraph()
    pg = wire.Graph()
    cam, dims = np.arange(8.0), np.array([10, 20], np.uint64)
    z = (np.zeros((0, 2)), np.zeros(0, np.float32), np.zeros((0, 8), np.uint64))
    a = [rg.add_node("a", np.zeros(7), 0, cam, dims, *z, 0), rg.add_node("b", np.zeros(7), 0, cam, dims, *z, 0)]
    b = [pg.add_node(cam, dims, *z, path="a", pose=np.zeros(7)), pg.add_node(cam, dims, *z, path="b", pose=np.zeros(7))]
    assert a == b
    e = ((np.zeros(0, np.uint64),) * 2 + (np.zeros(0),), np.zeros((0, 4)), np.zeros((0, 3), np.uint64), 2,
         np.full(9, np.nan), np.full(32, np.nan))
    for _ in range(5):
        assert rg.add_edge(a[0], a[1], *e) == pg.add_edge(b[0], b[1], *e)
    # reread: the generator restarts (a deserialized MeasurementGraph has a fresh one); edge ids only avoid other
    # EDGE ids (graph.hpp:88-93), so the first new edge gets the first draw again -- the first node's id
    text = pg.serialize()
    assert text == rg.serialize()
    pg2 = wire.Graph(text)
    assert pg2.add_edge(b[0], b[1], *e) == b[0] and pg2.add_edge(b[0], b[1], *e) == b[1]
    third = pg2.add_edge(b[0], b[1], *e)  # draws 3..7 are taken by the five edges: skipped
    assert pg2.num_edges == 8 and len({pg2.edge(i)["id"] for i in range(8)}) == 8
    rg3 = ref_io.RefGraph()  # the 8th draw of a fresh reference graph
    ids = [rg3.add_node("n", np.zeros(7), 0, cam, dims, *z, 0) for _ in range(8)]
    assert third == ids[7]


def test_throughput_of_reader_and_writer(hostlib):
    """Not a benchmark assertion, a sanity floor: a 2 000-feature node parses and writes at >= 20 MB/s."""
    import time
    rng = np.random.default_rng(2)
    g = wire.Graph()
    n = 20000
    rows = rng.integers(0, 2 ** 64, (n, 8), dtype=np.uint64)
    rows[:, 7] &= np.uint64((1 << 38) - 1)
    g.add_node(np.arange(8.0), np.array([4000, 3000], np.uint64), rng.uniform(0, 4000, (n, 2)),
               rng.uniform(0, 0.1, n).astype(np.float32), rows, node_id=1)
    t0 = time.perf_counter()
    text = g.serialize()
    t1 = time.perf_counter()
    g2 = wire.Graph(text)
    t2 = time.perf_counter()
    assert np.array_equal(g2.node(0)["rows"], rows)
    mb = len(text) / 1e6
    assert mb / (t1 - t0) > 20 and mb / (t2 - t1) > 20, (mb / (t1 - t0), mb / (t2 - t1))
