import sys, os, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/oracle")
import numpy as np
import oc_oracle as O
from opencalibration_b200 import capi
capi.init(0)
o = O.Oracle()
corr, _ = o.scene_homography(1500, 500, 3)
rng = np.random.default_rng(0)
for h in (32, 1024, 8192, 65536):
    samples = np.stack([rng.choice(len(corr), 4, replace=False) for _ in range(h)]).astype(np.uint32)
    capi.fit_homography(corr, samples)
    t0 = time.perf_counter()
    for _ in range(5):
        capi.fit_homography(corr, samples)
    print(h, (time.perf_counter() - t0) / 5 * 1e3, "ms per call (incl. copies)")
