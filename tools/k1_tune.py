"""K1 tuning helper (run under gpurun): times variants on the 10k x 10k pair, device resident.
usage: k1_tune.py [--variants 0,3,9] [--items 16] [--reps 20] [--col 0|1|2(both)]"""
import argparse, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
from opencalibration_b200 import capi, synthetic
import oc_oracle as O

ap = argparse.ArgumentParser()
ap.add_argument("--variants", default="0,1,2,3,4,5,6,7,8,9")
ap.add_argument("--items", default="16")
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--col", type=int, default=2)
ap.add_argument("--n", type=int, default=10000)
ap.add_argument("--check", type=int, default=1)
ap.add_argument("--update", default="0")
args = ap.parse_args()
capi.init(0)
n = args.n
a, b = synthetic.config2_pair(n, n, seed=1)
dq = torch.from_numpy(a.view(np.int64)).cuda(); dc = torch.from_numpy(b.view(np.int64)).cuda()
dout = torch.zeros(n, dtype=torch.int64, device="cuda"); dcol = torch.zeros(n, dtype=torch.int32, device="cuda")
wsb = capi.match_top2_workspace_bytes(n, n, True)
ws = torch.zeros(wsb + 4096, dtype=torch.uint8, device="cuda"); wsp = (ws.data_ptr() + 255) // 256 * 256
st = torch.cuda.current_stream().cuda_stream
if args.check:
    o = O.Oracle(); bk, bd, sd = o.match_top2(a[:256], b); cb = o.match_col_best(a, b[:128])
for upd, items in [(int(u), int(x)) for u in args.update.split(",") for x in args.items.split(",")]:
    capi.set_option("k1_items_per_sm", items)
    capi.set_option("k1_update", upd)
    for v in [int(x) for x in args.variants.split(",")]:
        capi.set_option("k1_variant", v)
        line = f"variant {v:2d} items/SM {items:3d} upd {upd}:"
        for col in ([False, True] if args.col == 2 else [bool(args.col)]):
            dcp = dcol.data_ptr() if col else None
            for _ in range(3):
                capi.match_top2_device(dq.data_ptr(), n, dc.data_ptr(), n, dout.data_ptr(), dcp, wsp, wsb, st)
            torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.reps):
                capi.match_top2_device(dq.data_ptr(), n, dc.data_ptr(), n, dout.data_ptr(), dcp, wsp, wsb, st)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.reps
            ok = ""
            if args.check:
                r = dout.cpu().numpy().view(capi.TOP2_DTYPE)
                good = np.array_equal(r["best_k"][:256], bk) and np.array_equal(r["best_d"][:256], bd) and np.array_equal(r["second_d"][:256], sd)
                if col:
                    good = good and np.array_equal(dcol.cpu().numpy().view(np.uint32)[:128], cb)
                ok = f" ok={good}"
            line += f"  {'col' if col else 'fwd'} {ms:.4f} ms {n*n/ms/1e6:6.1f} Gcmp/s{ok}"
        print(line, flush=True)
