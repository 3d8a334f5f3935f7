"""Runs on the GPU box: the forms of the tensor-core search kernel (option k1t_variant) on the configs[1] pair.

For every form: results compared record by record with the integer-pipe engine (K1, itself pinned to the oracle by the
tests), device-resident time per step with the L2 flushed between steps (CUDA events on the launching stream), and the
end-to-end rate through match_features_subset(std::vector<feature_2d>...) for several numbers of concurrent callers.

    python tools/k1t_tune.py [--variants 1,2,3] [--callers 1,4,8,12] [--steps 30] > gpurun_out/k1t_tune.json
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variants", default="1,2,3")
    ap.add_argument("--callers", default="1,4,8,12,16")
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--n", type=int, default=10000)
    args = ap.parse_args()
    import torch
    from opencalibration_b200 import capi, host, synthetic
    capi.init(0)
    n1 = n2 = args.n
    a, b = synthetic.config2_pair(n1, n2, seed=1)
    dq = torch.from_numpy(a.view(np.int64)).cuda()
    dc = torch.from_numpy(b.view(np.int64)).cuda()
    dout = torch.zeros(n1, dtype=torch.int64, device="cuda")
    dcol = torch.zeros(n2, dtype=torch.int32, device="cuda")
    wsb = capi.match_top2_workspace_bytes(n1, n2, True)
    ws = torch.zeros(wsb + 512, dtype=torch.uint8, device="cuda")
    wsp = (ws.data_ptr() + 255) // 256 * 256
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream

    def step(col=True):
        capi.match_top2_device(dq.data_ptr(), n1, dc.data_ptr(), n2, dout.data_ptr(), dcol.data_ptr() if col else None,
                               wsp, wsb, stream)

    def run(engine, variant, col=True):
        capi.set_option("k1_engine", engine)
        capi.set_option("k1t_variant", variant)
        dout.zero_()
        dcol.zero_()
        for _ in range(3):
            step(col)
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        for e0, e1 in ev:
            flush.fill_(1)
            e0.record()
            step(col)
            e1.record()
        torch.cuda.synchronize()
        ms = sorted(e0.elapsed_time(e1) for e0, e1 in ev)
        return (sum(ms) / len(ms), ms[len(ms) // 2], dout.cpu().numpy().view(capi.TOP2_DTYPE).copy(),
                dcol.cpu().numpy().copy())

    out = {"n1": n1, "n2": n2, "forms": {}}
    ms_int, med_int, want, want_col = run(1, 0)
    out["integer_pipes_ms"] = ms_int
    fa, fb = host.FeatureSet(a), host.FeatureSet(b)
    callers = [int(c) for c in args.callers.split(",") if c]
    for v in [int(x) for x in args.variants.split(",") if x]:
        rec = {}
        try:
            ms, med, got, got_col = run(2, v)
            rec["ms_mean"], rec["ms_median"] = ms, med
            rec["Gcmp_per_s"] = n1 * n2 / (ms * 1e-3) / 1e9
            rec["equal_to_integer_engine"] = bool(
                np.array_equal(got["best_k"], want["best_k"]) and np.array_equal(got["best_d"], want["best_d"]) and
                np.array_equal(got["second_d"], want["second_d"]) and np.array_equal(got_col, want_col))
            ms_nc, _, got_nc, _ = run(2, v, col=False)
            rec["ms_without_cross_check"] = ms_nc
            rec["equal_without_cross_check"] = bool(np.array_equal(got_nc, want))
            e2e = {}
            for c in callers:
                c = max(1, min(c, os.cpu_count() or 1))
                host.run_parallel_handles([fa] * c, [fb] * c, threads=c, cross_check=True, reps=2)
                reps = max(2, 96 // c)
                secs, _ = host.run_parallel_handles([fa] * c, [fb] * c, threads=c, cross_check=True, reps=reps)
                e2e[str(c)] = {"ms_per_call": secs * 1e3 / (reps * c),
                               "Gcmp_per_s": n1 * n2 * reps * c / secs / 1e9}
            rec["e2e_by_callers"] = e2e
        except Exception as e:  # noqa: BLE001
            rec["error"] = repr(e)
            out["forms"][str(v)] = rec
            print(json.dumps(out), flush=True)
            raise
        out["forms"][str(v)] = rec
    capi.set_option("k1_engine", 0)
    capi.set_option("k1t_variant", 0)
    out["host_cores"] = os.cpu_count()
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
