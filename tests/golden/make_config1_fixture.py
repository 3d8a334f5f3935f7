"""Generates tests/golden/config1_features.npz -- BASELINE.json configs[0]: the repo's own aerial pair.

Run in the build container only (needs /root/reference/test/test_data and Python cv2). Restates the reference's
extraction so that the fixture holds exactly what test/test_match.cpp:10-24 feeds to the hot path:
  * src/extract/extract_features.cpp:14-36: grayscale, INTER_AREA resize to max side 1600, AKAZE MLDB-486 (3 channels,
    threshold 5e-5); location = pt / scale, strength = response, descriptor = the 61 bytes bit-packed LSB first;
  * :55-87: strength-sorted non-maximum suppression at 8 px (in scaled pixels) -- the "sparse" features, to which
    test_match.cpp:17-18 truncates.
Both the oracle and the GPU path consume this dump, so the OpenCV version used here does not affect parity.
"""
import os
import sys

import cv2
import numpy as np

REF = "/root/reference/test/test_data"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "config1_features.npz")


def extract_sparse(path):
    im = cv2.imread(path)
    g = cv2.cvtColor(im, cv2.COLOR_BGR2GRAY)
    h, w = g.shape
    scale = float(min(np.float32(1.0), np.float32(1600) / np.float32(max(w, h))))
    g = cv2.resize(g, (0, 0), fx=scale, fy=scale, interpolation=cv2.INTER_AREA)
    akaze = cv2.AKAZE_create(cv2.AKAZE_DESCRIPTOR_MLDB, 486, 3, 0.00005)
    kps, desc = akaze.detectAndCompute(g, None)
    n = len(kps)
    xy = np.array([[np.float32(k.pt[0]) / scale, np.float32(k.pt[1]) / scale] for k in kps], np.float64)
    strength = np.array([k.response for k in kps], np.float32)
    rows = np.zeros((n, 64), np.uint8)
    rows[:, :61] = desc
    assert (rows[:, 60] >> 6).max() == 0  # bits 486..487 of the last byte are zero
    # std::sort by strength descending (ties: order is irrelevant to the fixture's purpose)
    order = np.argsort(-strength, kind="stable")
    xy, strength, rows = xy[order], strength[order], rows[order]
    # NMS: keep iff squared distance to the nearest kept point * scale^2 > 8^2
    lim = 64.0 / (scale * scale)
    cell = np.sqrt(lim)
    grid = {}
    keep = []

    def key(p):
        return (int(np.floor(p[0] / cell)), int(np.floor(p[1] / cell)))

    for i in range(n):
        p = xy[i]
        if i == 0:
            keep.append(i)
            grid.setdefault(key(p), []).append(i)
            continue
        cx, cy = key(p)
        ok = True
        for gx in range(cx - 2, cx + 3):
            for gy in range(cy - 2, cy + 3):
                for j in grid.get((gx, gy), ()):
                    d = p - xy[j]
                    if not (d[0] * d[0] + d[1] * d[1] > lim):
                        ok = False
                        break
                if not ok:
                    break
            if not ok:
                break
        if ok:
            keep.append(i)
            grid.setdefault(key(p), []).append(i)
    keep = np.array(keep)
    return xy[keep], strength[keep], rows[keep].view(np.uint64).reshape(-1, 8), n


def main():
    out = {}
    for tag, name in (("a", "P2540254.JPG"), ("b", "P2530253.JPG")):
        xy, st, desc, n_all = extract_sparse(os.path.join(REF, name))
        print(name, "keypoints", n_all, "sparse", len(xy))
        out[f"{tag}_xy"], out[f"{tag}_strength"], out[f"{tag}_desc"] = xy, st, desc
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    sys.exit(main())
