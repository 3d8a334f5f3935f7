"""Generates tests/golden/wire_vectors.npz and tests/golden/wire_graph.json with the reference's OWN graph.json writer
(src/io/serialize_MeasurementGraph.cpp + base64.c, compiled in place into oracle/_ref/liboc_ref_io.so against the
rapidjson headers of this image -- see oracle/Makefile target ref_io). Runs only in the build container
(/root/reference is needed); the fixtures travel.   python tests/golden/make_wire_vectors.py
"""
import os
import re
import struct
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oc_ref_io as R  # noqa: E402

CAM = np.array([3000.0, 2000.5, 1500.25, -0.1, 0.01, 1e-4, 1e-5, -2e-5])
DIMS = np.array([4000, 3000], np.uint64)


def special_doubles():
    v = [0.0, -0.0, 1.0, -1.0, 0.1, 0.2, 0.3, 1 / 3, 2 / 3, 1e21, 1e22, 9.999999999999999e20, 123456789012345678901.0,
         1e-5, 1e-6, 1e-7, 1.5e-7, 0.000001234, 5e-324, 2.2250738585072014e-308, 2.225073858507201e-308,
         1.7976931348623157e308, float("nan"), float("inf"), float("-inf"), 2.0 ** 53, 2.0 ** 53 + 2, 2.0 ** 63,
         2.0 ** 64, 1e15, 1e16, 1e17, 123456.789, 0.005, 0.01, 486.0, 1 / 486, 243 / 486, 0.8, 4.35, 0.000035,
         1e100, 1e-100, 1.2345678901234567e-300, 9007199254740993.0, 0.1 + 0.2, 100.0, 1e23, 8.41e21, 5e-5]
    v += [k / 486 for k in range(0, 487, 7)]  # match distances are multiples of 1/486
    v += [2.0 ** e for e in range(-1074, 1024, 37)]
    return np.array(v)


def make_doubles(seed=11, n_random=3000):
    rng = np.random.default_rng(seed)
    bits = rng.integers(0, 2 ** 64, n_random, dtype=np.uint64)
    bits = bits[(bits >> np.uint64(52)) & np.uint64(0x7FF) != np.uint64(0x7FF)]  # finite
    rand = bits.view(np.float64)
    pixels = np.round(rng.uniform(0, 6000, 1500), 3)                       # keypoint-like coordinates
    widened = rng.uniform(0, 0.02, 1500).astype(np.float32).astype(np.float64)  # float strengths widened
    unit = rng.normal(size=1000)                                           # homography / pose entries
    d = np.concatenate([special_doubles(), rand, pixels, widened, unit])
    if len(d) % 2:
        d = np.append(d, 1.0)
    return d


def doubles_through_reference(d):
    g = R.RefGraph()
    xy = d.reshape(-1, 2)
    g.add_node("d", np.zeros(7), 0, CAM, DIMS, xy, np.zeros(len(xy), np.float32), np.zeros((len(xy), 8), np.uint64),
               len(xy))
    text = g.serialize().decode()
    found = re.findall(r'"location": \[([^,\]]+), ([^,\]]+)\]', text)
    assert len(found) == len(xy)
    return [t for pair in found for t in pair]


def small_graph(seed=5):
    """3 image nodes, 3 edges (one without inliers, NaN poses), written by the reference."""
    rng = np.random.default_rng(seed)
    g = R.RefGraph()
    ids = []
    for k in range(3):
        n = 40 + 7 * k
        rows = rng.integers(0, 2 ** 64, (n, 8), dtype=np.uint64)
        rows[:, 7] &= np.uint64((1 << 38) - 1)
        xy = np.round(rng.uniform(0, 4000, (n, 2)), 2)
        st = rng.uniform(0.0005, 0.02, n).astype(np.float32)
        pose = np.concatenate([rng.normal(size=3) * 100, [0.0, 0.0, np.sin(0.3 * k), np.cos(0.3 * k)]])
        strings = dict(make="ACME \"Aerial\"", model="X\\1\t", serial_no="sn-%d" % k, lens_make="", lens_model="ü-lens",
                       datum="WGS-84", timestamp="12:0%d:00" % k, datestamp="2020:01:0%d" % (k + 1))
        cap = np.array([47.1 + k * 1e-4, 8.5, 500.0, 80.0, 0.5, -89.5, 12.25, 0.5, 1.0])
        ids.append(g.add_node("/data/img_%d.JPG" % k, pose, k % 2, CAM * (1 + 0.01 * (k % 2)), DIMS, xy, st, rows, n - 5,
                              strings, cap if k else None))
    for (a, b, nm, ni) in [(0, 1, 12, 6), (1, 0, 9, 0), (2, 1, 5, 5)]:
        i1, i2 = rng.integers(0, 40, nm).astype(np.uint64), rng.integers(0, 40, nm).astype(np.uint64)
        dist = np.sort(rng.integers(20, 200, nm))[::-1] / 486
        px = np.round(rng.uniform(0, 4000, (ni, 4)), 2)
        ix = np.stack([i1[:ni], i2[:ni], np.arange(ni, dtype=np.uint64)], axis=1) if ni else np.zeros((0, 3), np.uint64)
        H = np.eye(3) + rng.normal(size=(3, 3)) * 1e-3
        poses = np.full((4, 8), np.nan)
        poses[:, 0] = [ni, 0, 0, 0]
        if ni:
            poses[0, 1:] = [0.0, 0.0, 0.1, 0.995, 0.5, -0.25, 1e-9]
            poses[1, 1:] = [0.0, 0.0, -0.1, 0.995, -0.5, 0.25, -1e-9]
        g.add_edge(ids[a], ids[b], (i1, i2, dist), px, ix, 0 if ni else 2, H, poses)
    return g.serialize()


def base64_vectors(seed=3):
    rng = np.random.default_rng(seed)
    plain = [rng.integers(0, 256, n, dtype=np.uint8).tobytes() for n in list(range(0, 70)) + [61, 61, 183, 1000]]
    coded = [R.base64_encode(p) for p in plain]
    # decoder behaviour on text the encoder never produces (base64.c:127-170)
    odd = [b"", b"A", b"AB", b"ABC", b"ABCD", b"ABCDE", b"AB=D", b"A=", b"ABC=", b"AB==EF", b"AB CD", b"AB\nCD",
           b"ABCD!EFG", b"-_-_", b"QUJD" * 5 + b"Q", b"QUJD" * 5 + b"QU", b"=ABC"]
    odd_plain = [R.base64_decode(t) for t in odd]
    return plain, coded, odd, odd_plain


def main():
    assert R.build(), "oracle/_ref/liboc_ref_io.so could not be built here"
    d = make_doubles()
    texts = doubles_through_reference(d)
    plain, coded, odd, odd_plain = base64_vectors()
    pack = lambda items: (np.frombuffer(b"".join(items), np.uint8),  # noqa: E731
                          np.cumsum([0] + [len(i) for i in items]).astype(np.int64))
    out = {"double_bits": d.view(np.uint64)}
    for name, items in [("double_text", [t.encode() for t in texts]), ("b64_plain", plain), ("b64_coded", coded),
                        ("b64_odd", odd), ("b64_odd_plain", odd_plain)]:
        out[name + "_bytes"], out[name + "_offsets"] = pack(items)
    np.savez_compressed(os.path.join(HERE, "wire_vectors.npz"), **out)
    with open(os.path.join(HERE, "wire_graph.json"), "wb") as fh:
        fh.write(small_graph())
    print("doubles", len(d), "base64", len(plain) + len(odd), "graph bytes",
          os.path.getsize(os.path.join(HERE, "wire_graph.json")))


if __name__ == "__main__":
    main()
