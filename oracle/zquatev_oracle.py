"""CPU oracle for the zquatev hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``zquatev_b200/`` imports this module; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may.  Two things live here:

1. ``ref_*``  -- a numpy restatement of the REFERENCE algorithm, one function per
   reference routine, each citing the file:line it follows
   (``/root/reference/unblocked.cc``, ``zquatev.cc``, ``test.cc``).  The blocked
   routine (``blocked.cc:45-546``) applies exactly the same sequence of unitary
   symplectic transformations as ``unblocked.cc:44-131`` (it only delays them in
   compact-WY form), so the column-by-column restatement is the specification of
   both.
2. ``RefLib`` -- a ctypes door onto ``oracle/_ref/libzquatev_ref.so``: the
   UNMODIFIED reference compiled in place from ``/root/reference`` (recipe:
   ``oracle/Makefile``).  It pins the restatement (tests/test_oracle.py) and is the
   CPU baseline of ``bench.py``.

Parity pinning: the reference holds no golden vectors (``test.cc`` prints, never
asserts).  The restatement is pinned against (a) the reference built here, (b) the
golden eigenvalues of SURVEY.md Appendix B (``tests/golden/testcc_eigs.json``,
generated from the reference by ``tests/golden/make_golden.py``).
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(_HERE, "_ref", "libzquatev_ref.so")
REF_TEST_X = os.path.join(_HERE, "_ref", "test_ref.x")


# --------------------------------------------------------------------------------------
# inputs
# --------------------------------------------------------------------------------------
def gen_testcc(n: int):
    """The reference's test matrix, bit for bit (test.cc:58-78).

    glibc ``srand(32)``; for i in 0..n-1, j in 0..i four draws ``rand()%10000*1e-4``.
    Returns (A, B, C) with C the full 2n x 2n matrix ``[[A, B], [-conj(B), conj(A)]]``
    (column-major semantics: C[row, col]).
    """
    libc = ctypes.CDLL("libc.so.6")
    libc.srand(32)
    rnd = libc.rand
    A = np.zeros((n, n), dtype=np.complex128)
    B = np.zeros((n, n), dtype=np.complex128)
    for i in range(n):
        for j in range(i + 1):
            t0 = rnd() % 10000 * 0.0001
            t1 = rnd() % 10000 * 0.0001
            t2 = rnd() % 10000 * 0.0001
            t3 = rnd() % 10000 * 0.0001
            # test.cc:64-67 -- element (row j, col i) of a column-major n x n array
            A[j, i] = t0 if i == j else complex(t0, t1)
            A[i, j] = np.conj(A[j, i])
            B[j, i] = 0.0 if i == j else complex(t2, t3)
            B[i, j] = -B[j, i]
    C = np.zeros((2 * n, 2 * n), dtype=np.complex128)
    # test.cc:71-78
    C[:n, :n] = A
    C[n:, n:] = np.conj(A)
    C[:n, n:] = B
    C[n:, :n] = -np.conj(B)
    return A, B, C


def gen_sym(n: int, seed: int):
    """G_sym(n, seed): random quaternion-Hermitian matrix with entries in [-1/2, 1/2).

    D Hermitian, E antisymmetric; returns the full 2n x 2n matrix
    ``[[D, -conj(E)], [E, conj(D)]]`` (the layout zquatev.h:40-46 documents).
    """
    rng = np.random.default_rng(seed)
    X = rng.random((n, n)) - 0.5 + 1j * (rng.random((n, n)) - 0.5)
    Y = rng.random((n, n)) - 0.5 + 1j * (rng.random((n, n)) - 0.5)
    D = np.tril(X, -1)
    D = D + D.conj().T + np.diag(X.diagonal().real)
    E = np.tril(Y, -1)
    E = E - E.T
    return assemble(D, E)


def gen_spectrum(n: int, lam, seed: int = 0):
    """G_spec: quaternion-Hermitian matrix with prescribed (doubly degenerate) spectrum
    ``lam`` (length n), built as Q diag(lam) Q^H with Q a product of random quaternion
    reflectors (unitary symplectic), so clusters / exact repeats stress deflation."""
    rng = np.random.default_rng(seed)
    lam = np.asarray(lam, dtype=np.float64)
    D = np.diag(lam).astype(np.complex128)
    E = np.zeros((n, n), dtype=np.complex128)
    M = assemble(D, E)
    for _ in range(4):
        va = rng.standard_normal(n) + 1j * rng.standard_normal(n)
        vb = rng.standard_normal(n) + 1j * rng.standard_normal(n)
        P = np.zeros((2 * n, 2), dtype=np.complex128)
        P[:n, 0], P[n:, 0] = va, vb
        P[:n, 1], P[n:, 1] = -np.conj(vb), np.conj(va)
        P /= np.sqrt((abs(va) ** 2 + abs(vb) ** 2).sum())
        H = np.eye(2 * n) - 2.0 * P @ P.conj().T
        M = H @ M @ H.conj().T
    Dn, En = M[:n, :n], M[n:, :n]
    Dn = 0.5 * (Dn + Dn.conj().T)
    En = 0.5 * (En - En.T)
    return assemble(Dn, En)


def assemble(D, E):
    """Phi(D, E) = [[D, -conj(E)], [E, conj(D)]] (zquatev.h:40-44)."""
    n = D.shape[0]
    M = np.empty((2 * n, 2 * n), dtype=np.complex128)
    M[:n, :n] = D
    M[n:, :n] = E
    M[:n, n:] = -np.conj(E)
    M[n:, n:] = np.conj(D)
    return M


# --------------------------------------------------------------------------------------
# LAPACK scalar kernels the reference calls (f77.h:68-69), restated
# --------------------------------------------------------------------------------------
def _zlarfg(alpha, x):
    """LAPACK zlarfg: H^H [alpha; x] = [beta; 0], H = I - tau [1;v][1;v]^H, beta real.
    Called at unblocked.cc:58,104 / blocked.cc:176,345."""
    xnorm = np.linalg.norm(x)
    alphr, alphi = alpha.real, alpha.imag
    if xnorm == 0.0 and alphi == 0.0:
        return alpha, x.copy(), 0.0 + 0.0j
    beta = -np.copysign(np.sqrt(alphr * alphr + alphi * alphi + xnorm * xnorm), alphr)
    tau = complex((beta - alphr) / beta, -alphi / beta)
    v = x / (alpha - beta)
    return complex(beta), v, tau


def _zlartg(f, g):
    """LAPACK zlartg: [c s; -conj(s) c] [f; g] = [r; 0], c real.
    Called at unblocked.cc:86 / blocked.cc:237."""
    if g == 0:
        return 1.0, 0.0 + 0.0j, f
    if f == 0:
        r = abs(g)
        return 0.0, np.conj(g) / r, complex(r)
    f1 = abs(f)
    nrm = np.hypot(f1, abs(g))
    c = f1 / nrm
    s = (f / f1) * np.conj(g) / nrm
    r = (f / f1) * nrm
    return c, s, r


def _zrot(x, y, c, s):
    """BLAS zrot: x' = c x + s y ; y' = c y - conj(s) x (unblocked.cc:88,91,95)."""
    tx = c * x + s * y
    ty = c * y - np.conj(s) * x
    return tx, ty


# --------------------------------------------------------------------------------------
# the reference algorithm, restated
# --------------------------------------------------------------------------------------
def ref_unblocked_update(D0, D1, Q0, Q1):
    """Restatement of ts::impl::unblocked_update (unblocked.cc:44-131) on whole
    n x n arrays (in place).  On return D0 is Hermitian tridiagonal (last sub-diagonal
    possibly complex), D1 ~ 0, Phi(Q0, Q1) the accumulated transformation."""
    n = D0.shape[0]
    for k in range(n - 1):                                   # unblocked.cc:51
        ln = n - k - 1
        if ln > 1:                                           # H1, unblocked.cc:53-81
            alpha = D1[k + 1, k]
            _, v1, tau = _zlarfg(alpha, D1[k + 2:, k].copy())   # :58
            vec = np.concatenate(([1.0 + 0j], v1))
            tau = np.conj(tau)                               # :59
            cvec = np.conj(vec)                              # :61
            # 00 block (:64-67)
            tmp = D0[k + 1:, k:].conj().T @ cvec             # zgemv "C"
            tmp[1:] += (-np.conj(tau) * 0.5 * np.vdot(tmp[1:], cvec)) * cvec
            D0[k + 1:, k:] += -np.conj(tau) * np.outer(cvec, np.conj(tmp))   # zgerc
            D0[k:, k + 1:] += -tau * np.outer(tmp, vec)                      # zgeru
            # 10 block (:70-72)
            tmp = D1[k:, k + 1:] @ cvec                      # zgemv "N"
            D1[k + 1:, k:] += tau * np.outer(vec, tmp)
            D1[k:, k + 1:] += -tau * np.outer(tmp, vec)
            # Q update (:75-80)
            tmp = Q0[:, k + 1:] @ cvec
            Q0[:, k + 1:] += -tau * np.outer(tmp, vec)
            tmp = Q1[:, k + 1:] @ cvec
            Q1[:, k + 1:] += -tau * np.outer(tmp, vec)
        # symplectic Givens, unblocked.cc:84-96
        c, s, _ = _zlartg(D0[k + 1, k], D1[k + 1, k])
        D0[k + 1, k:], D1[k + 1, k:] = _zrot(D0[k + 1, k:].copy(), D1[k + 1, k:].copy(), c, s)
        x, y = _zrot(np.conj(D1[k:, k + 1]), D0[k:, k + 1].copy(), c, s)
        D1[k:, k + 1], D0[k:, k + 1] = np.conj(x), y
        x, y = _zrot(np.conj(Q1[:, k + 1]), Q0[:, k + 1].copy(), c, s)
        Q1[:, k + 1], Q0[:, k + 1] = np.conj(x), y
        if ln > 1:                                           # H2, unblocked.cc:99-128
            alpha = D0[k + 1, k]
            _, v1, tau = _zlarfg(alpha, D0[k + 2:, k].copy())   # :104
            vec = np.concatenate(([1.0 + 0j], v1))
            tau = np.conj(tau)
            cvec = np.conj(vec)
            # 00 (:110-113)
            tmp = D0[k + 1:, k:].conj().T @ vec
            tmp[1:] += (-tau * 0.5 * np.vdot(vec, tmp[1:])) * vec
            D0[k + 1:, k:] += -tau * np.outer(vec, np.conj(tmp))             # zgerc
            D0[k:, k + 1:] += -np.conj(tau) * np.outer(tmp, np.conj(vec))    # zgerc
            # 01 (:116-118)
            tmp = D1[k + 1:, k:].T @ vec                     # zgemv "T"
            D1[k + 1:, k:] += -np.conj(tau) * np.outer(cvec, tmp)            # zgeru
            D1[k:, k + 1:] += np.conj(tau) * np.outer(tmp, np.conj(vec))     # zgerc
            # Q (:121-126)
            tmp = Q0[:, k + 1:] @ vec
            Q0[:, k + 1:] += -np.conj(tau) * np.outer(tmp, np.conj(vec))
            tmp = -(Q1[:, k + 1:] @ vec)
            Q1[:, k + 1:] += np.conj(tau) * np.outer(tmp, np.conj(vec))


def ref_zquatev(Dfull):
    """Restatement of ts::zquatev (zquatev.cc:42-100) for ld2 == n2.

    ``Dfull`` is the (2n, 2n) complex matrix of which only the left half is read
    (zquatev.h:45-46).  Returns (eig[n] ascending, out[2n,2n] = (U,-V*;V,U*), info)."""
    n2 = Dfull.shape[0]
    assert n2 % 2 == 0                                       # zquatev.cc:43
    n = n2 // 2
    D0 = np.array(Dfull[:n, :n], dtype=np.complex128)        # repack, zquatev.cc:48-54
    D1 = np.array(Dfull[n:, :n], dtype=np.complex128)
    Q0 = np.eye(n, dtype=np.complex128)                      # zquatev.cc:57-61
    Q1 = np.zeros((n, n), dtype=np.complex128)
    ref_unblocked_update(D0, D1, Q0, Q1)                     # zquatev.cc:68-72
    # band pack + zhbev("V","L",kd=1), zquatev.cc:79-84: Hermitian tridiagonal, lower
    T = np.diag(D0.diagonal().real).astype(np.complex128)
    for i in range(n - 1):
        T[i + 1, i] = D0[i + 1, i]
        T[i, i + 1] = np.conj(D0[i + 1, i])
    info = 0
    if not np.all(np.isfinite(T)):
        return np.full(n, np.nan), np.full((n2, n2), np.nan, dtype=np.complex128), n - 1
    eig, Z = np.linalg.eigh(T)
    U = Q0 @ Z                                               # zquatev.cc:87
    V = Q1 @ Z                                               # zquatev.cc:88-90
    out = np.empty((n2, n2), dtype=np.complex128)
    out[:n, :n], out[n:, :n] = U, V
    out[:n, n:], out[n:, n:] = -np.conj(V), np.conj(U)       # zquatev.cc:93-98
    return eig, out, info


# --------------------------------------------------------------------------------------
# checks printed by the reference's own test (test.cc:104-112) + north_star metrics
# --------------------------------------------------------------------------------------
def testcc_checks(C, out, eig, eig_zheev=None):
    """``error`` = ||V^H M V - Lambda||_F^2 with Lambda_ii = eig[i % n] (test.cc:104-108) and
    ``maxdev`` = max_i |eig_zheev[2i] - eig[i]| (test.cc:110-112)."""
    n = C.shape[0] // 2
    R = out.conj().T @ (C @ out)
    R[np.diag_indices(2 * n)] -= np.concatenate([eig, eig])
    err = float(np.vdot(R, R).real)
    maxdev = None
    if eig_zheev is not None:
        maxdev = float(np.max(np.abs(eig_zheev[0::2] - eig)))
    return err, maxdev


def quality(M, out, eig):
    """north_star metrics (N = 2n, Frobenius norms, eps = 2^-52):
    residual ||M V - V L||/(N ||M|| eps), orthogonality ||V^H V - I||/(N eps),
    pairing = max |right half - Theta(left half)| (must be exactly 0)."""
    N = M.shape[0]
    n = N // 2
    eps = np.finfo(np.float64).eps
    lam = np.concatenate([eig, eig])
    res = np.linalg.norm(M @ out - out * lam[None, :]) / (N * np.linalg.norm(M) * eps)
    orth = np.linalg.norm(out.conj().T @ out - np.eye(N)) / (N * eps)
    U, V = out[:n, :n], out[n:, :n]
    pair = max(np.max(np.abs(out[:n, n:] + np.conj(V))), np.max(np.abs(out[n:, n:] - np.conj(U))))
    return float(res), float(orth), float(pair)


# --------------------------------------------------------------------------------------
# the reference itself, compiled here
# --------------------------------------------------------------------------------------
class RefLib:
    """ctypes binding of oracle/_ref/libzquatev_ref.so (built by oracle/Makefile from the
    unmodified /root/reference sources).  ``zquatev`` = ts::zquatev (zquatev.h:54)."""

    def __init__(self, path: str = REF_SO):
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} missing: run `make -C oracle` (needs /root/reference)")
        self.lib = ctypes.CDLL(path)
        self.lib.zq_ref_zquatev.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        self.lib.zq_ref_zquatev.restype = ctypes.c_int
        self.lib.zq_ref_zheev.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        self.lib.zq_ref_zheev.restype = ctypes.c_int
        self.lib.zq_ref_set_threads.argtypes = [ctypes.c_int]
        self.lib.zq_ref_get_threads.restype = ctypes.c_int

    @staticmethod
    def available(path: str = REF_SO) -> bool:
        return os.path.exists(path)

    def set_threads(self, t: int):
        self.lib.zq_ref_set_threads(int(t))

    def get_threads(self) -> int:
        return int(self.lib.zq_ref_get_threads())

    def zquatev(self, M):
        """M: (2n, 2n) complex array, M[row, col].  Returns (eig[n], out[2n,2n], info)."""
        n2 = M.shape[0]
        buf = np.asfortranarray(M, dtype=np.complex128).copy(order="F")
        eig = np.zeros(n2, dtype=np.float64)
        info = self.lib.zq_ref_zquatev(n2, buf.ctypes.data, n2, eig.ctypes.data)
        return eig[: n2 // 2].copy(), buf, int(info)

    def zheev(self, M):
        n2 = M.shape[0]
        buf = np.asfortranarray(M, dtype=np.complex128).copy(order="F")
        eig = np.zeros(n2, dtype=np.float64)
        info = self.lib.zq_ref_zheev(n2, buf.ctypes.data, n2, eig.ctypes.data)
        return eig, buf, int(info)
