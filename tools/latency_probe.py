"""Latency of the small RANSAC-scoring calls (ocb_score_bound, h=32 / h=1, n=2000) from 1..16 host threads, with and
without a large batched K1 submission running on the same GPU. Run under gpurun."""
import os, sys, threading, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from opencalibration_b200 import capi, synthetic
import ctypes as C
capi.init(0)
L = capi.lib()
L.ocb_corr_bind.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
L.ocb_score_bound.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
corr, H = synthetic.homography_scene(2000, 0, seed=1, noise=0.0005)
models = synthetic.random_models(0, 32, seed=3, base=H)
order = np.random.default_rng(0).permutation(2000).astype(np.uint32)

def worker(reps, out, idx, h):
    capi.init(0)
    capi.check(L.ocb_corr_bind(corr.ctypes.data, 2000, order.ctypes.data))
    score = np.zeros(32); count = np.zeros(32, np.uint32); bits = np.zeros((32, 63), np.uint32)
    t0 = time.perf_counter()
    for _ in range(reps):
        capi.check(L.ocb_score_bound(0, models.ctypes.data, h, 0.005, 1 if h > 1 else 0, score.ctypes.data, count.ctypes.data,
                                     bits.ctypes.data if h == 1 else None))
    out[idx] = (time.perf_counter() - t0) / reps

def run(threads, h, reps=300):
    out = [0] * threads
    ts = [threading.Thread(target=worker, args=(reps, out, i, h)) for i in range(threads)]
    [t.start() for t in ts]; [t.join() for t in ts]
    return np.mean(out) * 1e6

images, _pos, pairs = synthetic.grid_survey(4, 4, 8192, seed=7)
for i, im in enumerate(images):
    capi.register_descriptors(1000 + i, im)
plist = [(1000 + a, 1000 + b) for a, b in pairs]
nq = [8192] * len(plist)
res = np.zeros(len(plist) * 8192, capi.TOP2_DTYPE)
capi.match_pairs(plist, nq, out=res)
stop = False
def hog():
    capi.init(0)
    while not stop:
        capi.match_pairs(plist, nq, out=res)
for busy in (False, True):
    if busy:
        th = threading.Thread(target=hog); th.start(); time.sleep(0.2)
    for threads in (1, 4, 16):
        for h in (32, 1):
            print(f"K1 batch running={busy} threads={threads} h={h}: {run(threads, h):.0f} us per call", flush=True)
stop = True; th.join()
