"""Development probe: K1 (k_matvec) alone over a ladder of trailing sizes m, per-launch CUDA-event time and the
achieved fraction of the measured copy bandwidth.  The kernel variant is chosen by the environment (ZQ_K1_TPB, ...),
read once per process, so variants are compared by running this script once per setting.
usage: k1_sweep.py [n] [reps]"""
import ctypes
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zquatev_b200 import api  # noqa: E402
from bench import make_input, measured_peaks  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
dev = torch.device("cuda", 0)
A = make_input(n, dev)
v = torch.randn((n, 4), dtype=torch.float64, device=dev)
y = torch.zeros((n, 4), dtype=torch.float64, device=dev)
peak = measured_peaks()[0]["hbm_gbs"]
rows = []
for m in [16384, 14336, 12288, 10240, 8192, 6144, 5120, 4096, 3072, 2048, 1024, 512]:
    if m > n:
        continue
    s = n - m
    ms = ctypes.c_double(0)
    rc = api.lib().zq_test_matvec(n, s, A.data_ptr(), 2 * n, v.data_ptr(), y.data_ptr(), reps, ctypes.byref(ms))
    assert rc == 0, rc
    gbs = (16.0 * m * m + 64.0 * m) / (ms.value * 1e-3) * 1e-9
    rows.append({"m": m, "us": round(ms.value * 1e3, 2), "GBs": round(gbs, 1), "frac": round(gbs / peak, 4)})
print(json.dumps({"n": n, "env": {k: v for k, v in os.environ.items() if k.startswith("ZQ_")}, "peak": peak, "k1": rows}))
