"""Generates tests/golden/ref_vectors.npz by running the REFERENCE'S OWN object code (oracle/_ref: its
src/match/match_features.cpp and src/model_inliers/ransac.cpp compiled in place) on committed / seeded inputs.

Run in the build container only (needs /root/reference to build oracle/_ref). What each vector pins:
  c1_*   config 1 (tests/golden/config1_features.npz): spatially_subsample_feature_indices(.., 40.0) for both
         images and match_features_subset on them -- pure reference code, bit-exact pin for the match path;
  c2s_*  a small config-2-shaped synthetic pair (seeded, with planted ties) through match_features_subset;
  rs_*   ransac<H|E|F> through the reference's driver (sampling, SPRT, LO, termination) on seeded scenes and on the
         config-1 matches; the model member functions it calls are the oracle's restatement (Eigen is not available),
         so these pin the driver exactly and the model arithmetic only as restated.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oc_oracle as O  # noqa: E402
from opencalibration_b200 import synthetic  # noqa: E402


def unit_rays(xy, f=5000.0, cols=5344, rows=4016):
    """image_to_3d for an undistorted planar camera (src/distort/distort_keypoints.cpp:67-98) with the camera of
    test/test_ransac_functional.cpp:26-31."""
    p = (xy - np.array([cols, rows]) * 0.5) / f
    h = np.concatenate([p, np.ones((len(p), 1))], axis=1)
    return h / np.linalg.norm(h, axis=1, keepdims=True)


def correspondences(fa, fb, i1, i2, dist):
    c = np.zeros((len(i1), 7))
    c[:, 0:3] = unit_rays(fa[i1])
    c[:, 3:6] = unit_rays(fb[i2])
    c[:, 6] = dist
    return c


def main():
    O.build(ref=True)
    ref = O.Reference()
    orc = O.Oracle()
    out = {}
    fx = np.load(os.path.join(HERE, "config1_features.npz"))
    ia = ref.subsample(fx["a_xy"], fx["a_strength"], 40.0)
    ib = ref.subsample(fx["b_xy"], fx["b_strength"], 40.0)
    m1, m2, md = ref.match_features_subset(fx["a_desc"], fx["b_desc"], ia, ib)
    print("config1: idx", len(ia), len(ib), "matches", len(m1))
    out.update(c1_idx_a=ia, c1_idx_b=ib, c1_m1=m1, c1_m2=m2, c1_dist=md)
    corr = correspondences(fx["a_xy"], fx["b_xy"], m1, m2, md)
    s, M, inl = ref.ransac(O.KIND_H, corr)
    print("config1 ransac H: score", s, "inliers", int(inl.sum()))
    out.update(rs_c1_corr=corr, rs_c1_score=s, rs_c1_M=M, rs_c1_inl=inl)

    a, b = synthetic.config2_pair(1200, 1000, seed=3)
    b[500] = b[17]
    b[999] = b[17]  # planted exact ties
    rng = np.random.default_rng(11)
    i1 = rng.permutation(1200)[:900]
    i2 = rng.permutation(1000)[:800]
    s1, s2, sd = ref.match_features_subset(a, b, i1, i2)
    print("c2-small: matches", len(s1))
    out.update(c2s_a=a, c2s_b=b, c2s_i1=i1, c2s_i2=i2, c2s_m1=s1, c2s_m2=s2, c2s_dist=sd)

    scenes = {
        "h30": (O.KIND_H, orc.scene_homography(140, 60)[0]),
        "h80": (O.KIND_H, orc.scene_homography(40, 160)[0]),
        "hdeg": (O.KIND_H, orc.scene_homography_near_degenerate()[0]),
        "f30": (O.KIND_F, orc.scene_fundamental(140, 60)[0]),
        "fplane": (O.KIND_F, orc.scene_fundamental(200, 0, 0.8)[0]),
        "e0": (O.KIND_E, orc.scene_fundamental(200, 0)[0]),
    }
    # PROSAC branch: same H scene with qualities drawn from a seeded generator
    hq = orc.scene_homography(140, 60)[0].copy()
    hq[:, 6] = np.random.default_rng(43).uniform(0.05, 0.4, len(hq))
    scenes["h30q"] = (O.KIND_H, hq)
    for name, (kind, c) in scenes.items():
        s, M, inl = ref.ransac(kind, c)
        print(f"scene {name}: score {s:.6f} inliers {int(inl.sum())}")
        out[f"rs_{name}_kind"] = kind
        out[f"rs_{name}_corr"] = c
        out[f"rs_{name}_score"] = s
        out[f"rs_{name}_M"] = M
        out[f"rs_{name}_inl"] = inl
    path = os.path.join(HERE, "ref_vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
