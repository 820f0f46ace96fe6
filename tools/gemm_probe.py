"""Development probe: the K4 / K6 GEMM shapes of the 2n = 32768 solve, stacked complex (3M / 4-product) against the
quaternion 8-product kernel, per-launch CUDA-event time and canonical TFLOP/s (8 flop per complex multiply-add, 32 per
quaternion multiply-add).  usage: gemm_probe.py [n]"""
import ctypes
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zquatev_b200 import api  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
kb = 64
L = api.lib()
ms = ctypes.c_double(0)
rnd = lambda r, c: torch.rand((c, r, 2), dtype=torch.float64, device="cuda") - 0.5      # column-major r x c complex
out = []
for m in (n, n // 2):
    # K4: (D + jE)[m x m] -= [V W][W V]^H, lower
    Aq, Sq, C = rnd(2 * m, 2 * kb), rnd(2 * m, 2 * kb), rnd(2 * n, m)
    L.zq_test_qgemm(0, 1, m, m, 2 * kb, -1.0, Aq.data_ptr(), 2 * m, m, Sq.data_ptr(), 2 * m, m, 1.0, C.data_ptr(), 2 * n, n, 1, 5, ctypes.byref(ms))
    fl = 32.0 * m * m * 2 * kb / 2
    out.append({"op": "K4 q8", "m": m, "ms": round(ms.value, 3), "canonical_tflops": round(fl / ms.value * 1e-9, 2)})
    Lc, Rc = rnd(2 * m, 4 * kb), rnd(m, 4 * kb)
    al = (ctypes.c_double * 2)(-1.0, 0.0)
    be = (ctypes.c_double * 2)(1.0, 0.0)
    for three_m in (1, 0):
        L.zq_test_set_gemm_3m(three_m)
        L.zq_test_zgemm(0, 1, 2 * m, m, 4 * kb, al, Lc.data_ptr(), 2 * m, Rc.data_ptr(), m, be, C.data_ptr(), 2 * n, 0, 5, ctypes.byref(ms))
        fl = 8.0 * 2 * m * m * 4 * kb
        out.append({"op": "K4-shaped stacked complex (full, not lower) " + ("3M" if three_m else "4-product"), "m": m, "ms": round(ms.value, 3),
                    "canonical_tflops": round(fl / ms.value * 1e-9, 2)})
    del Aq, Sq, C, Lc, Rc
    # K6 update: X[m rows, n cols] -= V (m x kb) TY (kb x n)
    P, TY, X = rnd(2 * m, kb), rnd(2 * kb, n), rnd(2 * n, n)
    L.zq_test_qgemm(0, 0, m, n, kb, -1.0, P.data_ptr(), 2 * m, m, TY.data_ptr(), 2 * kb, kb, 1.0, X.data_ptr(), 2 * n, n, 0, 5, ctypes.byref(ms))
    fl = 32.0 * m * n * kb
    out.append({"op": "K6 update q8", "m": m, "ms": round(ms.value, 3), "canonical_tflops": round(fl / ms.value * 1e-9, 2)})
    # K6 Y = V^H X (no split-K through this door: the plain launch)
    Y = rnd(2 * kb, n)
    L.zq_test_qgemm(1, 0, kb, n, m, 1.0, P.data_ptr(), 2 * m, m, X.data_ptr(), 2 * n, n, 0.0, Y.data_ptr(), 2 * kb, kb, 0, 5, ctypes.byref(ms))
    out.append({"op": "K6 Y=V^H X q8 (no split-K)", "m": m, "ms": round(ms.value, 3), "canonical_tflops": round(fl / ms.value * 1e-9, 2)})
    # the same two products with two panels merged (K = 128 quaternions for the update, 128 rows of Y)
    P2, TY2, Y2 = rnd(2 * m, 2 * kb), rnd(4 * kb, n), rnd(4 * kb, n)
    L.zq_test_qgemm(0, 0, m, n, 2 * kb, -1.0, P2.data_ptr(), 2 * m, m, TY2.data_ptr(), 4 * kb, 2 * kb, 1.0, X.data_ptr(), 2 * n, n, 0, 5, ctypes.byref(ms))
    out.append({"op": "K6 update q8, two panels merged (K = 128)", "m": m, "ms": round(ms.value, 3), "canonical_tflops": round(2 * fl / ms.value * 1e-9, 2)})
    L.zq_test_qgemm(1, 0, 2 * kb, n, m, 1.0, P2.data_ptr(), 2 * m, m, X.data_ptr(), 2 * n, n, 0.0, Y2.data_ptr(), 4 * kb, 2 * kb, 0, 5, ctypes.byref(ms))
    out.append({"op": "K6 Y=V^H X q8, two panels merged (128 rows, no split-K)", "m": m, "ms": round(ms.value, 3), "canonical_tflops": round(2 * fl / ms.value * 1e-9, 2)})
    del P2, TY2, Y2
    L.zq_test_set_gemm_3m(1)
    Pc = rnd(2 * m, 2 * kb)
    al = (ctypes.c_double * 2)(-1.0, 0.0)
    L.zq_test_zgemm(0, 0, m, n, 2 * kb, al, Pc.data_ptr(), 2 * m, TY.data_ptr(), 2 * kb, be, X.data_ptr(), 2 * n, 0, 5, ctypes.byref(ms))
    out.append({"op": "K6 update stacked complex 3M (one of two halves)", "m": m, "ms": round(2 * ms.value, 3),
                "canonical_tflops": round(fl / (2 * ms.value) * 1e-9, 2)})
    del P, TY, X, Y, Pc
print(json.dumps({"n": n, "env": {k: v for k, v in os.environ.items() if k.startswith("ZQ_")}, "gemm": out}))
