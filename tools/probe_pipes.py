"""Prints the issue-rate probes of include/ocb_probe.h (lane-ops / clk / SM) as one JSON line. Run under gpurun."""
import ctypes as C, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opencalibration_b200 import capi
capi.init(0)
probe = np.zeros(32)
capi.check(capi.lib().ocb_probe_pipes(probe.ctypes.data_as(C.c_void_p), 32))
names = ["sms", "sm_mhz", "popc", "lop3", "imad", "iadd", "imnmx", "isetp_sel", "mix_popc_lop3", "mix_popc_lop3_imad",
         "dadd", "dmul", "dfma", "ddiv", "dsqrt", "dfma3"]
print(json.dumps({n: round(float(probe[i]), 3) for i, n in enumerate(names)}))
