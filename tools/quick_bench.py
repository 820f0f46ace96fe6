"""Development probe (not the contract bench): device-resident solves at a few sizes with the
phase split, K1 stand-alone bandwidth, GEMM rates.  Prints JSON lines."""
import ctypes
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zquatev_b200 as z  # noqa: E402
from zquatev_b200 import api  # noqa: E402


def rand_qh(n, seed=32):
    g = torch.Generator(device="cuda").manual_seed(seed)
    buf = torch.empty((2 * n, 2 * n), dtype=torch.complex128, device="cuda")   # buf[col, row]
    X = torch.rand((n, 2 * n, 2), dtype=torch.float64, device="cuda", generator=g) - 0.5
    buf[:n] = torch.view_as_complex(X)
    # symmetrise lower->upper not needed: only lower triangles are read; make diag of D real
    idx = torch.arange(n, device="cuda")
    buf[idx, idx] = buf[idx, idx].real.to(torch.complex128)
    return buf


def main():
    gemm_only = "gemm" in sys.argv[1:]
    sizes = [int(x) for x in sys.argv[1:] if x.isdigit()] or [512, 1024, 2048, 4096]
    if gemm_only:
        sizes = []
    for n in sizes:
        buf = rand_qh(n)
        eig = torch.zeros(n, dtype=torch.float64, device="cuda")
        for rep in range(2):
            b2 = buf.clone()
            torch.cuda.synchronize()
            t0 = time.time()
            info = z.zquatev_device(2 * n, b2.data_ptr(), 2 * n, eig.data_ptr())
            torch.cuda.synchronize()
            dt = time.time() - t0
        ph = z.last_phases()
        flops = 164.0 / 3.0 * n ** 3
        print(json.dumps({"n": n, "info": info, "wall_s": dt, "tflops_canonical": flops / dt * 1e-12, "phases_ms": ph}),
              flush=True)
        del b2
    # K1 stand-alone
    for n in ([] if gemm_only else [4096, 8192, 16384]):
        A = torch.rand((n, 2 * n, 2), dtype=torch.float64, device="cuda")
        v = torch.rand((n, 4), dtype=torch.float64, device="cuda")
        y = torch.zeros((n, 4), dtype=torch.float64, device="cuda")
        ms = ctypes.c_double(0)
        rc = api.lib().zq_test_matvec(n, 0, A.data_ptr(), 2 * n, v.data_ptr(), y.data_ptr(), 10, ctypes.byref(ms))
        byts = 16.0 * n * n
        print(json.dumps({"k1_n": n, "rc": rc, "ms": ms.value, "GBps_lower": byts / ms.value * 1e-6}), flush=True)
        del A
    # GEMM shapes
    for (ta, tb, M, N, K, lower, name) in [(0, 1, 8192, 8192, 128, 1, "trailing nb32"), (0, 1, 8192, 8192, 256, 1, "trailing nb64"),
                                           (1, 0, 64, 8192, 8192, 0, "Y=P^H X"), (0, 0, 8192, 8192, 64, 0, "X-=P TY nb32"),
                                           (0, 0, 8192, 8192, 128, 0, "X-=P TY nb64"), (0, 0, 4096, 4096, 4096, 0, "square"),
                                           (1, 0, 128, 8192, 8192, 0, "Y=P^H X nb64"), (0, 0, 16384, 16384, 64, 0, "X-=P TY nb32 16k")]:
        cr = lambda r, c: torch.rand((c, r, 2), dtype=torch.float64, device="cuda")
        A = cr(K, M) if ta else cr(M, K)
        B = cr(N, K) if tb else cr(K, N)
        C = cr(M, N)
        al = (ctypes.c_double * 2)(-1.0, 0.0)
        be = (ctypes.c_double * 2)(1.0, 0.0)
        ms = ctypes.c_double(0)
        lda = K if ta else M
        ldb = N if tb else K
        rc = api.lib().zq_test_zgemm(ta, tb, M, N, K, al, A.data_ptr(), lda, B.data_ptr(), ldb, be, C.data_ptr(), M, lower, 3,
                                     ctypes.byref(ms))
        fl = 8.0 * M * N * K * (0.5 if lower else 1.0)
        print(json.dumps({"gemm": name, "rc": rc, "ms": ms.value, "tflops": fl / ms.value * 1e-9}), flush=True)
    # cuBLAS reference points (NOT used by the product; context for the FP64 roofline)
    for dt_, nm in ([] if gemm_only else [(torch.float64, "cublas_dgemm"), (torch.complex128, "cublas_zgemm")]):
        a = torch.rand((4096, 4096), dtype=torch.float64, device="cuda").to(dt_)
        b = torch.rand((4096, 4096), dtype=torch.float64, device="cuda").to(dt_)
        torch.matmul(a, b)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(5):
            torch.matmul(a, b)
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / 5
        fl = (2.0 if dt_ == torch.float64 else 8.0) * 4096 ** 3
        print(json.dumps({"ref": nm, "ms": ms, "tflops": fl / ms * 1e-9}), flush=True)


if __name__ == "__main__":
    main()
