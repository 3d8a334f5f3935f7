"""K2 tuning helper (run under gpurun): times the scoring kernel variants on configs[2] (4096 hypotheses x 20000
correspondences, device resident) and checks a slice against the oracle.
usage: k2_tune.py [--variants 0,1] [--hg 0,8,7,6] [--reps 10]"""
import argparse, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
from opencalibration_b200 import capi, host, synthetic
import oc_oracle as O

ap = argparse.ArgumentParser()
ap.add_argument("--variants", default="1,2")
ap.add_argument("--hg", default="0,8,7,6")
ap.add_argument("--reps", type=int, default=10)
args = ap.parse_args()
capi.init(0)
H, N = 4096, 20000
st = torch.cuda.current_stream().cuda_stream
corr, _ = synthetic.homography_scene(14000, 6000, seed=42)
d_corr7 = torch.from_numpy(corr).cuda()
d_corr4 = torch.zeros(N * 4, dtype=torch.float64, device="cuda")
order = np.random.default_rng(0).permutation(N).astype(np.uint32)
d_order = torch.from_numpy(order.view(np.int32)).cuda()
d_pos = torch.zeros(N, dtype=torch.int32, device="cuda")
capi.prepare_correspondences_device(d_corr7.data_ptr(), d_order.data_ptr(), N, d_corr4.data_ptr(), d_pos.data_ptr(), st)
d_score = torch.zeros(H, dtype=torch.float64, device="cuda")
d_count = torch.zeros(H, dtype=torch.int32, device="cuda")
orc = O.Oracle()
for kind, name, thr in ((0, "homography", 0.005), (1, "epipolar", 0.01)):
    if kind == 0:
        rng = np.random.default_rng(5)
        models = np.stack([host.fit(0, corr, rng.choice(N, 4, replace=False).astype(np.uintp)) for _ in range(H)])
    else:
        models = synthetic.random_models(kind, H, seed=3)
    d_models = torch.from_numpy(np.ascontiguousarray(models)).cuda()
    so, co, _ = orc.score_hypotheses(kind, models[:24], corr, order=order, thr=thr)
    for v in [int(x) for x in args.variants.split(",")]:
        for hg in [int(x) for x in args.hg.split(",")]:
            capi.set_option("k2_variant", v); capi.set_option("k2_hg", hg)
            fn = lambda: capi.score_models_device(kind, d_models.data_ptr(), H, d_corr4.data_ptr(), None, N, thr,
                                                  d_score.data_ptr(), d_count.data_ptr(), None, st)
            for _ in range(2):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.reps):
                fn()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.reps
            ok = np.array_equal(d_score.cpu().numpy()[:24], so) and np.array_equal(d_count.cpu().numpy().view(np.uint32)[:24], co)
            print(f"{name:10s} variant {v} hg {hg}: {ms:.4f} ms  {H*N/ms/1e3:9.0f} M residuals/s  ok={ok}", flush=True)
