"""Runs on the GPU box: cycles per tcgen05.mma.kind::i8 (M 128 x N x K 32) by N, number of independent accumulators
and location of A (include/ocb_probe.h, ocb_probe_umma). Prints one JSON object."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opencalibration_b200 import capi  # noqa: E402

capi.init(0)
out = {"what": "cycles per MMA, M 128 x N x K 32 s8, one issuing thread per SM, 148 SMs", "rows": []}
for ts in (0, 1):
    for n in (64, 128, 256):
        for chains in (1, 2, 3, 4, 6, 8):
            if chains * n > (384 if ts else 512):
                continue
            for ctas in (1, 148):
                cyc = capi.probe_umma(ts, n, chains, 512, ctas)
                out["rows"].append({"a_in_tmem": ts, "n": n, "chains": chains, "ctas": ctas, "cycles_per_mma": round(cyc, 2),
                                    "arithmetic_cycles": n / 2})
print(json.dumps(out))
