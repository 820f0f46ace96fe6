"""helpers for the -m gpu tests: torch is only the device-memory / stream plumbing."""
import ctypes

import numpy as np
import torch

from zquatev_b200 import api


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def fdev(a):
    """Fortran-ordered (column-major) numpy array -> device tensor holding the same memory."""
    a = np.asfortranarray(a)
    return torch.from_numpy(a.T.copy()).cuda()        # C-order of the transpose == F-order of a


def fhost(t, shape):
    """device tensor with column-major memory of `shape` -> numpy (row, col) array"""
    return t.cpu().numpy().reshape(shape[::-1]).T


def left_half(M):
    """(2n x n) column-major [D; E] block = left half of M."""
    n = M.shape[0] // 2
    return np.asfortranarray(M[:, :n])


def quats(a, b):
    """planar complex pair -> (n, 4) doubles a.re a.im b.re b.im"""
    return np.stack([a.real, a.imag, b.real, b.imag], axis=1).astype(np.float64).copy()


def matvec(M, s, va, vb, reps=0):
    n = M.shape[0] // 2
    A = fdev(left_half(M))
    v = dev(quats(va, vb))
    y = torch.zeros((n, 4), dtype=torch.float64, device="cuda")
    ms = ctypes.c_double(0)
    rc = api.lib().zq_test_matvec(n, s, A.data_ptr(), 2 * n, v.data_ptr(), y.data_ptr(), reps, ctypes.byref(ms))
    assert rc == 0, rc
    yh = y.cpu().numpy()
    return yh[:, 0] + 1j * yh[:, 1], yh[:, 2] + 1j * yh[:, 3], ms.value


def zgemm(ta, tb, alpha, A, B, beta, C, lower=0, reps=0):
    """A, B, C numpy (row, col) arrays as STORED (before op)."""
    M, N = C.shape
    K = A.shape[0] if ta else A.shape[1]
    Ad, Bd, Cd = fdev(A), fdev(B), fdev(C)
    al = (ctypes.c_double * 2)(alpha.real, alpha.imag)
    be = (ctypes.c_double * 2)(beta.real, beta.imag)
    ms = ctypes.c_double(0)
    rc = api.lib().zq_test_zgemm(ta, tb, M, N, K, al, Ad.data_ptr(), A.shape[0], Bd.data_ptr(), B.shape[0], be,
                                 Cd.data_ptr(), C.shape[0], lower, reps, ctypes.byref(ms))
    assert rc == 0, rc
    return fhost(Cd, C.shape), ms.value


def stedc(d, e):
    n = len(d)
    dd = dev(np.asarray(d, dtype=np.float64))
    ee = dev(np.concatenate([np.asarray(e, dtype=np.float64), [0.0]]))
    w = torch.zeros(n, dtype=torch.float64, device="cuda")
    Z = torch.zeros((n, n), dtype=torch.float64, device="cuda")
    rc = api.lib().zq_test_stedc(n, dd.data_ptr(), ee.data_ptr(), w.data_ptr(), Z.data_ptr())
    return rc, w.cpu().numpy(), Z.cpu().numpy().T


def bisect(d, e):
    n = len(d)
    dd = dev(np.asarray(d, dtype=np.float64))
    ee = dev(np.concatenate([np.asarray(e, dtype=np.float64), [0.0]]))
    w = torch.zeros(n, dtype=torch.float64, device="cuda")
    rc = api.lib().zq_test_bisect(n, dd.data_ptr(), ee.data_ptr(), w.data_ptr())
    assert rc == 0
    return w.cpu().numpy()


def tridiag(M, nb):
    n = M.shape[0] // 2
    A = fdev(left_half(M))
    d = torch.zeros(n, dtype=torch.float64, device="cuda")
    e = torch.zeros(n, dtype=torch.float64, device="cuda")
    tau = torch.zeros(n, dtype=torch.float64, device="cuda")
    al = torch.zeros((n, 4), dtype=torch.float64, device="cuda")
    rc = api.lib().zq_test_tridiag(n, nb, A.data_ptr(), 2 * n, d.data_ptr(), e.data_ptr(), tau.data_ptr(), al.data_ptr())
    assert rc == 0, rc
    Ah = fhost(A, (2 * n, n))
    return d.cpu().numpy(), e.cpu().numpy(), tau.cpu().numpy(), al.cpu().numpy(), Ah


def solve_host(M, jobz=1, nb=0, ld2=None):
    """ts::zquatev mirror on a host array; returns (eig, out, info)."""
    import zquatev_b200 as z
    n2 = M.shape[0]
    ld2 = ld2 or n2
    buf = np.zeros((ld2, n2), dtype=np.complex128, order="F")
    buf[:n2, :] = M
    eig = np.full(n2, -777.0)
    info = z.zquatev(n2, buf, ld2, eig, jobz=jobz, nb=nb)
    return eig, buf[:n2, :], info
