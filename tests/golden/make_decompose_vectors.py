"""Generates tests/golden/decompose_h.npz: golden vectors for homography_model::decompose
(reference src/model_inliers/homography_model.cpp:138-185).

  H, k, R, t, n   seeded homographies (general, pure rotations, near-rotations, scaled) and what the REAL
                  cv2.decomposeHomographyMat(H, I) of this container's OpenCV (the library the reference calls at
                  :146) returns for them -- pins the restated closed form;
  sort_scores,    every 4-tuple of scores in {-1,0,1,2,3} and the permutation libstdc++'s std::stable_sort produces
  sort_perm       with the reference's non-strict comparator `p1.score >= p2.score` (:180-181), from a 20-line C++
                  program compiled here with the same g++ -- pins the oracle's restatement of that sort.
Run in the build container (needs cv2 and g++).
"""
import itertools
import os
import subprocess
import tempfile

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

SORT_CPP = r"""
#include <algorithm>
#include <array>
#include <cstdio>
struct P { int score; int id; };
int main() {
    int s[4];
    while (std::scanf("%d %d %d %d", &s[0], &s[1], &s[2], &s[3]) == 4) {
        std::array<P, 4> p;
        for (int i = 0; i < 4; i++) p[i] = P{s[i], i};
        std::stable_sort(p.begin(), p.end(), [](const P &a, const P &b) { return a.score >= b.score; });
        std::printf("%d %d %d %d\n", p[0].id, p[1].id, p[2].id, p[3].id);
    }
}
"""


def main():
    rng = np.random.default_rng(2024)
    Hs = []
    for t in range(240):
        ang = rng.normal(0, 0.3, 3)
        Rm, _ = cv2.Rodrigues(ang)
        tt = rng.normal(0, 0.3, 3)
        nn = rng.normal(0, 1, 3)
        nn /= np.linalg.norm(nn)
        H = Rm + np.outer(tt, nn) / (1 + rng.uniform(0, 2))
        if t % 16 == 0:
            H = Rm.copy()                     # pure rotation -> single solution
        elif t % 16 == 1:
            H = Rm + 1e-5 * np.outer(tt, nn)  # just below the 0.001 rotation test
        elif t % 16 == 2:
            H = H * rng.uniform(0.1, 10)      # arbitrary scale (the reference normalises to H(2,2) = 1)
        else:
            H = H / H[2, 2]
        Hs.append(H)
    Hs.append(np.eye(3))
    Hs = np.array(Hs)
    k = np.zeros(len(Hs), np.int32)
    R = np.full((len(Hs), 4, 3, 3), np.nan)
    T = np.full((len(Hs), 4, 3), np.nan)
    N = np.full((len(Hs), 4, 3), np.nan)
    for i, H in enumerate(Hs):
        kk, Rs, Ts, Ns = cv2.decomposeHomographyMat(H, np.eye(3))
        k[i] = kk
        for j in range(kk):
            R[i, j], T[i, j], N[i, j] = Rs[j], Ts[j].ravel(), Ns[j].ravel()
    print("solutions histogram:", np.bincount(k), "cv2", cv2.__version__)

    combos = np.array(list(itertools.product([-1, 0, 1, 2, 3], repeat=4)), np.int32)
    with tempfile.TemporaryDirectory() as d:
        src, exe = os.path.join(d, "s.cpp"), os.path.join(d, "s")
        open(src, "w").write(SORT_CPP)
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-o", exe, src])
        out = subprocess.run([exe], input="\n".join(" ".join(map(str, c)) for c in combos), text=True,
                             capture_output=True, check=True).stdout
    perm = np.array([[int(x) for x in line.split()] for line in out.strip().splitlines()], np.int32)
    assert perm.shape == combos.shape
    np.savez_compressed(os.path.join(HERE, "decompose_h.npz"), H=Hs, k=k, R=R, t=T, n=N, sort_scores=combos,
                        sort_perm=perm, cv2_version=np.array(cv2.__version__))
    print("wrote decompose_h.npz:", len(Hs), "homographies,", len(combos), "sort cases")


if __name__ == "__main__":
    main()
