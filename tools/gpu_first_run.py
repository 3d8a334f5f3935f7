"""First-contact GPU script: pipe probes, K1/K2 parity vs the oracle, K1 variant sweep. Run under gpurun."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
from opencalibration_b200 import capi, synthetic
import oc_oracle as O
import ctypes as C

out = {}
capi.init(0)
L = capi.lib()
probe = np.zeros(16)
capi.check(L.ocb_probe_pipes(probe.ctypes.data_as(C.c_void_p), 16))
names = ["sms", "sm_mhz", "popc", "lop3", "imad", "iadd", "imnmx", "isetp_sel", "mix_popc_lop3", "mix_popc_lop3_imad",
         "dadd", "dmul", "dfma", "ddiv", "dsqrt"]
out["probe"] = {n: float(probe[i]) for i, n in enumerate(names)}
print("PROBE", json.dumps(out["probe"]), flush=True)

o = O.Oracle()
# ---- K1 parity on small shapes, all variants
rng = np.random.default_rng(5)
ok_all = True
for v in range(10):
    capi.set_option("k1_variant", v)
    for (n1, n2) in [(1, 1), (5, 3), (130, 64), (513, 1000), (1000, 65), (700, 1)]:
        a, b = synthetic.config2_pair(n1, n2, seed=n1 * 7 + n2)
        # duplicates for tie coverage
        if n2 > 4:
            b[n2 // 2] = b[1]; b[n2 - 1] = b[1]
        r, col = capi.match_top2(a, b, cross_check=True)
        bk, bd, sd = o.match_top2(a, b)
        cb = o.match_col_best(a, b)
        ok = np.array_equal(r["best_k"], bk) and np.array_equal(r["best_d"], bd) and np.array_equal(r["second_d"], sd) and np.array_equal(col, cb)
        ok_all &= ok
        if not ok:
            print("K1 MISMATCH variant", v, n1, n2, flush=True)
print("K1 small parity all variants:", ok_all, flush=True)
out["k1_small_parity"] = bool(ok_all)

# ---- K1 sweep on 10k x 10k device resident
a, b = synthetic.config2_pair(10000, 10000, seed=1)
t0 = time.time(); bk, bd, sd = o.match_top2(a[:512], b); t_or = time.time() - t0
dq = torch.from_numpy(a.view(np.int64)).cuda(); dc = torch.from_numpy(b.view(np.int64)).cuda()
dout = torch.zeros(10000, dtype=torch.int64, device="cuda")
ws_bytes = capi.match_top2_workspace_bytes(10000, 10000, False)
ws = torch.zeros(ws_bytes + 4096, dtype=torch.uint8, device="cuda")
wsp = (ws.data_ptr() + 255) // 256 * 256
st = torch.cuda.current_stream().cuda_stream
res = {}
for items in (8, 16, 24):
    capi.set_option("k1_items_per_sm", items)
    for v in range(10):
        capi.set_option("k1_variant", v)
        for _ in range(3):
            capi.match_top2_device(dq.data_ptr(), 10000, dc.data_ptr(), 10000, dout.data_ptr(), None, wsp, ws_bytes, st)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            capi.match_top2_device(dq.data_ptr(), 10000, dc.data_ptr(), 10000, dout.data_ptr(), None, wsp, ws_bytes, st)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        r = dout.cpu().numpy().view(capi.TOP2_DTYPE)
        good = np.array_equal(r["best_k"][:512], bk) and np.array_equal(r["best_d"][:512], bd) and np.array_equal(r["second_d"][:512], sd)
        dcol = torch.zeros(10000, dtype=torch.int32, device="cuda")
        wsb2 = capi.match_top2_workspace_bytes(10000, 10000, True)
        ws2 = torch.zeros(wsb2 + 4096, dtype=torch.uint8, device="cuda"); wsp2 = (ws2.data_ptr() + 255) // 256 * 256
        for _ in range(3):
            capi.match_top2_device(dq.data_ptr(), 10000, dc.data_ptr(), 10000, dout.data_ptr(), dcol.data_ptr(), wsp2, wsb2, st)
        torch.cuda.synchronize(); e0.record()
        for _ in range(10):
            capi.match_top2_device(dq.data_ptr(), 10000, dc.data_ptr(), 10000, dout.data_ptr(), dcol.data_ptr(), wsp2, wsb2, st)
        e1.record(); torch.cuda.synchronize()
        msc = e0.elapsed_time(e1) / 10
        res[f"v{v}_i{items}"] = dict(ms=ms, gcmp=1e8 / ms / 1e6, ok=bool(good), ms_col=msc)
        print(f"K1 variant {v} items/SM {items}: {ms:.4f} ms  {1e8/ms/1e6:.1f} Gcmp/s ok={good} | with cross-check {msc:.4f} ms {1e8/msc/1e6:.1f} Gcmp/s", flush=True)
out["k1_sweep"] = res

# ---- K2 parity + timing
for kind in (0, 2):
    corr, H = synthetic.homography_scene(700, 300, seed=9, noise=0.002)
    models = synthetic.random_models(kind, 37, seed=4, base=H)
    if kind == 0:
        models[0, :9] = H.T.ravel(); models[0, 9:] = np.linalg.inv(H).T.ravel()
    thr = 0.005 if kind == 0 else 0.01
    order = np.random.default_rng(2).permutation(len(corr)).astype(np.uint32)
    for od in (None, order):
        s, c, bits = capi.score_models(kind, models, corr, thr, order=od)
        so, co, bo = o.score_hypotheses(kind, models, corr, order=None if od is None else od.astype(np.uintp), thr=thr)
        print(f"K2 kind {kind} order={'yes' if od is not None else 'no'}: score exact={np.array_equal(s, so)} count={np.array_equal(c, co)} bits={np.array_equal(bits, bo)} maxinl={c.max()}", flush=True)
    e = capi.residuals(kind, models[0], corr)
    eo = np.array([o.error(kind, models[0], corr[i]) for i in range(len(corr))])
    print(f"K2 residuals kind {kind} exact={np.array_equal(e, eo)}", flush=True)
corr, H = synthetic.homography_scene(14000, 6000, seed=42)
models = synthetic.random_models(0, 4096, seed=4, base=H)
order = np.random.default_rng(2).permutation(len(corr)).astype(np.uint32)
for rep in range(3):
    t0 = time.time(); s, c, _ = capi.score_models(0, models, corr, 0.005, order=order, want_bits=False); t1 = time.time() - t0
print(f"K2 4096x20000 e2e host call: {t1*1e3:.2f} ms -> {4096*20000/t1/1e9:.2f} G residuals/s", flush=True)
out["k2_e2e_ms"] = t1 * 1e3
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "first_run.json"), "w"), indent=1)
