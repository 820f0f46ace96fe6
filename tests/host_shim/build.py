"""TEST INFRASTRUCTURE ONLY: builds tests/host_shim/_build/libdc_host.so (g++, no GPU) -- the
device D&C numerical cores instantiated for the host -- and wraps them for the oracle flow."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.path.join(HERE, "_build")
SO = os.path.join(OUT, "libdc_host.so")


def build():
    os.makedirs(OUT, exist_ok=True)
    src = os.path.join(HERE, "dc_host_shim.cc")
    hdr = os.path.join(ROOT, "zquatev_b200", "csrc", "dc_core.cuh")
    if os.path.exists(SO) and os.path.getmtime(SO) > max(os.path.getmtime(src), os.path.getmtime(hdr)):
        return SO
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-ffp-contract=off", "-I" + cuda_inc,
                           "-I" + os.path.join(ROOT, "zquatev_b200", "csrc"), src, "-o", SO])
    return SO


class HostCores:
    def __init__(self):
        self.lib = ctypes.CDLL(build())

    def leaf(self, d, e):
        m = len(d)
        dd = np.array(d, dtype=np.float64)
        ee = np.zeros(m, dtype=np.float64)
        ee[: m - 1] = e
        Z = np.eye(m, dtype=np.float64, order="F").copy(order="F")
        info = self.lib.zqh_leaf(m, dd.ctypes.data_as(ctypes.c_void_p), ee.ctypes.data_as(ctypes.c_void_p),
                                 Z.ctypes.data_as(ctypes.c_void_p), m)
        assert info == 0
        return dd, Z

    def secular(self, dl, z2, rho):
        k = len(dl)
        dl = np.ascontiguousarray(dl, dtype=np.float64)
        z2 = np.ascontiguousarray(z2, dtype=np.float64)
        org = np.zeros(k, dtype=np.int32)
        mu = np.zeros(k, dtype=np.float64)
        it = np.zeros(k, dtype=np.int32)
        P = ctypes.c_void_p
        worst = self.lib.zqh_secular(k, dl.ctypes.data_as(P), z2.ctypes.data_as(P), ctypes.c_double(rho),
                                     org.ctypes.data_as(P), mu.ctypes.data_as(P), it.ctypes.data_as(P))
        return org, mu, worst

    def deflate(self, rho, ds, zs, col, n1=None):
        nm = len(ds)
        ds = np.ascontiguousarray(ds, dtype=np.float64)
        zs = np.ascontiguousarray(zs, dtype=np.float64)
        col = np.ascontiguousarray(col, dtype=np.int32)
        dlam = np.zeros(nm); wz = np.zeros(nm); dfval = np.zeros(nm); rcc = np.zeros(nm); rss = np.zeros(nm)
        ndcol = np.zeros(nm, dtype=np.int32); dfcol = np.zeros(nm, dtype=np.int32); ndtype = np.zeros(nm, dtype=np.int32)
        n1 = nm // 2 if n1 is None else n1
        rc1 = np.zeros(nm, dtype=np.int32); rc2 = np.zeros(nm, dtype=np.int32)
        out3 = np.zeros(3, dtype=np.int32)
        P = ctypes.c_void_p
        a = lambda x: x.ctypes.data_as(P)
        self.lib.zqh_deflate(nm, n1, ctypes.c_double(rho), a(ds), a(zs), a(col), a(dlam), a(wz), a(ndcol), a(ndtype), a(dfval),
                             a(dfcol), a(rc1), a(rc2), a(rcc), a(rss), a(out3))
        k, nd, nr = (int(v) for v in out3)
        assert k + nd == nm
        self.last_types = ndtype[:k].copy()
        rots = [(int(rc1[i]), int(rc2[i]), float(rcc[i]), float(rss[i])) for i in range(nr)]
        return k, dlam[:k].copy(), wz[:k].copy(), ndcol[:k].copy(), dfval[:nd].copy(), dfcol[:nd].copy(), rots
