"""helpers for the -m gpu tests: torch is only the device-memory / stream plumbing."""
import ctypes
import math

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


def qgemm(ta, tb, alpha, Aa, Ab, Ba, Bb, beta, Ca, Cb, lower=0, reps=0):
    """quaternion GEMM door: operands are (a, b) complex pairs AS STORED (before op); each pair is packed into one
    column-major array with the b-part stacked below the a-part (b offset = number of rows of the a-part)."""
    M, N = Ca.shape
    K = Aa.shape[0] if ta else Aa.shape[1]
    A = np.vstack([Aa, Ab]); B = np.vstack([Ba, Bb]); C = np.vstack([Ca, Cb])
    Ad, Bd, Cd = fdev(A), fdev(B), fdev(C)
    ms = ctypes.c_double(0)
    rc = api.lib().zq_test_qgemm(ta, tb, M, N, K, float(alpha), Ad.data_ptr(), A.shape[0], Aa.shape[0], Bd.data_ptr(), B.shape[0], Ba.shape[0],
                                 float(beta), Cd.data_ptr(), C.shape[0], M, lower, reps, ctypes.byref(ms))
    assert rc == 0, rc
    out = fhost(Cd, C.shape)
    return out[:M], out[M:], ms.value


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


# ------------------------------------------------------------------------------------------------
# north_star quality metrics at sizes where the CPU oracle's O(N^3) numpy products are too slow:
# the SAME formulas as oracle.zquatev_oracle.quality (which restates test.cc:104-112), evaluated on the
# device with torch (cuBLAS zgemm) -- checker only, never on the product path.
# ------------------------------------------------------------------------------------------------
def build_DE(left0):
    """(D, E) as full n x n (row, col) tensors from the lower triangles of the column-major left half [n][2n]:
    M = [[D, -conj E], [E, conj D]], D Hermitian with real diagonal, E antisymmetric (zquatev.h:40-44)."""
    n = left0.shape[0]
    Dl = left0[:, :n].T                                   # (row, col) view; lower triangle valid
    El = left0[:, n:].T
    D = torch.tril(Dl, -1)
    D = D + D.conj().T
    D.diagonal().copy_(torch.diagonal(Dl).real.to(torch.complex128))
    E = torch.tril(El, -1)
    E = E - E.T
    return D, E


def device_quality(left0, out, eig, col_chunk=2048, rank=0, world=1, reduce=None, DE=None):
    """left0 : [n][2n] complex128 device tensor = column-major INPUT left half (A; B); only its lower
               triangles are meaningful (what the solver reads).
       out   : [2n][2n] complex128 device tensor = column-major RESULT (all 2n columns).
       eig   : [n] float64 device tensor.
    With world > 1 every rank checks the column chunks c = rank, rank + world, ... of ITS OWN copy of the
    gathered result and `reduce(t)` (a SUM all-reduce of a float64 tensor) combines the squared norms, so the
    numbers cover all 2n columns on every N and also see a wrong gather.
    Returns dict(residual, orthogonality, pairing, trace_err, sumsq_relerr, ascending, fro_norm):
       residual      = ||M V - V L||_F / (N ||M||_F eps),   orthogonality = ||V^H V - I||_F / (N eps)   (N = 2n)
       pairing       = max |right half - Theta(left half)|  (must be exactly 0)."""
    n = left0.shape[0]
    N = 2 * n
    eps = 2.220446049250313e-16
    dev_ = left0.device
    D, E = DE if DE is not None else build_DE(left0)
    Dl = left0[:, :n].T
    fro2 = 2.0 * (torch.linalg.norm(D) ** 2 + torch.linalg.norm(E) ** 2)
    lam = torch.cat([eig, eig])
    r2 = torch.zeros((), dtype=torch.float64, device=dev_)
    o2 = torch.zeros((), dtype=torch.float64, device=dev_)
    chunks = list(range(0, N, col_chunk))
    for ci, c0 in enumerate(chunks):
        if ci % world != rank:
            continue
        c1 = min(N, c0 + col_chunk)
        X = out[c0:c1].T                                  # N x nc, columns contiguous in memory
        Xa, Xb = X[:n], X[n:]
        # M X = [D Xa - conj(E) Xb ; E Xa + conj(D) Xb],  conj(E) Xb = conj(E conj(Xb))
        top = D @ Xa - (E @ Xb.conj()).conj()
        bot = E @ Xa + (D @ Xb.conj()).conj()            # conj(D) Xb = conj(D conj(Xb))
        l = lam[c0:c1][None, :]
        r2 += torch.linalg.norm(top - Xa * l) ** 2 + torch.linalg.norm(bot - Xb * l) ** 2
        del top, bot
        G = out @ X.conj()                                # conj of the N x nc block of V^H V (out as a matrix is V^T); no 2n x 2n temporary
        idx = torch.arange(c0, c1, device=dev_)
        G[idx, idx - c0] -= 1.0
        o2 += torch.linalg.norm(G) ** 2
        del G
    if reduce is not None:
        t = torch.stack([r2, o2])
        reduce(t)
        r2, o2 = t[0], t[1]
    # pairing: columns n + c = Theta(column c) = (-conj V_c ; conj U_c), exact
    pair = 0.0
    for c0 in range(0, n, col_chunk):
        c1 = min(n, c0 + col_chunk)
        L, R = out[c0:c1], out[n + c0:n + c1]             # [col][row]
        pair = max(pair, (R[:, :n] + L[:, n:].conj()).abs().max().item(), (R[:, n:] - L[:, :n].conj()).abs().max().item())
    fro = math.sqrt(fro2.item())
    tr = torch.diagonal(Dl).real.sum().item()
    return {"residual": math.sqrt(r2.item()) / (N * fro * eps), "orthogonality": math.sqrt(o2.item()) / (N * eps),
            "pairing": pair, "trace_err": abs(eig.sum().item() - tr),
            "sumsq_relerr": abs((eig ** 2).sum().item() - 0.5 * fro2.item()) / (0.5 * fro2.item()),   # ||M||_F^2 = 2 sum(lambda^2)
            "ascending": bool(torch.all(eig[1:] >= eig[:-1]).item()) if n > 1 else True, "fro_norm": fro,
            "eig_absmax": eig.abs().max().item()}
