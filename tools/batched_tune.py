"""Batched K1 (ocb_match_pairs) timing on a grid slice, per update form. Run under gpurun."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from opencalibration_b200 import capi, synthetic
capi.init(0)
rows, cols, n_desc = 6, 8, 8192
images, _pos, pairs = synthetic.grid_survey(rows, cols, n_desc, seed=7)
for i, im in enumerate(images):
    capi.register_descriptors(1000 + i, im)
plist = [(1000 + a, 1000 + b) for a, b in pairs]
nq = [n_desc] * len(plist)
res = np.zeros(len(plist) * n_desc, capi.TOP2_DTYPE)
for upd in (0, 1, 2):
    for variant in (0, 7, 8):
        capi.set_option("k1_update", upd); capi.set_option("k1_variant", variant)
        capi.match_pairs(plist[:8], nq[:8], out=res)
        for npairs in (64, 256, len(plist)):
            t0 = time.perf_counter(); capi.match_pairs(plist[:npairs], nq[:npairs], out=res); s = time.perf_counter() - t0
            print(f"update {upd} variant {variant} pairs {npairs}: {s*1e3:.1f} ms {npairs*n_desc*n_desc/s/1e9:.1f} Gcmp/s", flush=True)
