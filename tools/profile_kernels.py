"""(timings printed here are meaningless under ncu)
Launches the two dominant kernels once at headline-workload shapes (for `ncu --set full`):
K1 mat-vec on a 16384^2 trailing matrix and the DMMA GEMM at the trailing-update / back-transform
shapes of 2n = 32768 (nb = 64)."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zquatev_b200 import api

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
A = torch.rand((n, 2 * n, 2), dtype=torch.float64, device="cuda")
v = torch.rand((n, 4), dtype=torch.float64, device="cuda")
y = torch.zeros((n, 4), dtype=torch.float64, device="cuda")
ms = ctypes.c_double(0)
api.lib().zq_test_matvec(n, 0, A.data_ptr(), 2 * n, v.data_ptr(), y.data_ptr(), 2, ctypes.byref(ms))
print("k1 ms", ms.value, "GB/s", 16.0 * n * n / ms.value * 1e-6)
del A
m = n // 2
for (ta, tb, M, N, K, lower, name) in [(0, 1, m, m, 256, 1, "trailing"), (1, 0, 128, n, m, 0, "Y=P^H X"), (0, 0, m, n, 128, 0, "X-=P TY")]:
    cr = lambda r, c: torch.rand((c, r, 2), dtype=torch.float64, device="cuda")
    Aa = cr(K, M) if ta else cr(M, K)
    Bb = cr(N, K) if tb else cr(K, N)
    C = cr(M, N)
    al = (ctypes.c_double * 2)(-1.0, 0.0)
    be = (ctypes.c_double * 2)(1.0, 0.0)
    api.lib().zq_test_zgemm(ta, tb, M, N, K, al, Aa.data_ptr(), K if ta else M, Bb.data_ptr(), N if tb else K, be, C.data_ptr(), M, lower, 1,
                            ctypes.byref(ms))
    fl = 8.0 * M * N * K * (0.5 if lower else 1.0)
    print(name, "ms", ms.value, "TF/s", fl / ms.value * 1e-9)
# K5 last (one-CTA reduction of a 2n = 512 problem): a solve below n = 1024 switches the GEMMs to the 4-product kernel
import zquatev_b200 as z
from bench import make_input
ns = 256
ws = torch.empty((2 * ns, 2 * ns), dtype=torch.complex128, device="cuda")
ws[:ns].copy_(make_input(ns, torch.device("cuda", 0)))
es = torch.zeros(ns, dtype=torch.float64, device="cuda")
print("k5 solve info", z.zquatev_device(2 * ns, ws.data_ptr(), 2 * ns, es.data_ptr(), sync=True), z.last_phases())
