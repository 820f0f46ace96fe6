"""(timings printed here are meaningless under ncu)  Launches the quaternion 8-product GEMM (qgemm.cu) once at the three
shapes of the 2n = 32768 solve (m = 8192): K4 trailing update, K6 Y = V^H X (plain launch), K6 X -= V TY; and the 3M
complex kernel at the K6 update shape for comparison."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zquatev_b200 import api  # noqa: E402

n, m, kb = 16384, 8192, 64
L = api.lib()
ms = ctypes.c_double(0)
rnd = lambda r, c: torch.rand((c, r, 2), dtype=torch.float64, device="cuda") - 0.5
Aq, Sq, C = rnd(2 * m, 2 * kb), rnd(2 * m, 2 * kb), rnd(2 * n, m)
L.zq_test_qgemm(0, 1, m, m, 2 * kb, -1.0, Aq.data_ptr(), 2 * m, m, Sq.data_ptr(), 2 * m, m, 1.0, C.data_ptr(), 2 * n, n, 1, 0, ctypes.byref(ms))
del Aq, Sq, C
P, TY, X, Y = rnd(2 * m, kb), rnd(2 * kb, n), rnd(2 * n, n), rnd(2 * kb, n)
L.zq_test_qgemm(1, 0, kb, n, m, 1.0, P.data_ptr(), 2 * m, m, X.data_ptr(), 2 * n, n, 0.0, Y.data_ptr(), 2 * kb, kb, 0, 0, ctypes.byref(ms))
L.zq_test_qgemm(0, 0, m, n, kb, -1.0, P.data_ptr(), 2 * m, m, TY.data_ptr(), 2 * kb, kb, 1.0, X.data_ptr(), 2 * n, n, 0, 0, ctypes.byref(ms))
L.zq_test_set_gemm_3m(1)
Pc = rnd(2 * m, 2 * kb)
al = (ctypes.c_double * 2)(-1.0, 0.0)
be = (ctypes.c_double * 2)(1.0, 0.0)
L.zq_test_zgemm(0, 0, m, n, 2 * kb, al, Pc.data_ptr(), 2 * m, TY.data_ptr(), 2 * kb, be, X.data_ptr(), 2 * n, 0, 0, ctypes.byref(ms))
torch.cuda.synchronize()
print("done")
