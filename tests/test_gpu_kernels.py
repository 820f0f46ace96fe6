"""-m gpu: every CUDA kernel against its CPU restatement in oracle/ (through the C ABI test doors)."""
import numpy as np
import pytest

from oracle import quat_kernels as K
from oracle import tridiag_dc as T
from oracle import zquatev_oracle as O
from tests.test_oracle import _tri_cases

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,s", [(1, 0), (5, 0), (5, 2), (64, 0), (64, 63), (130, 0), (130, 7), (257, 129), (700, 0),
                                 (700, 333), (1031, 65)])
def test_k1_matvec(n, s):
    from tests import gpu_util as G
    M = O.gen_sym(n, 100 + n)
    rng = np.random.default_rng(n)
    va = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    vb = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    ya, yb, _ = G.matvec(M, s, va, vb)
    D, E = M[:n, :n], M[n:, :n]
    ra, rb = K.matvec_lower(D[s:, s:], E[s:, s:], va[s:], vb[s:])
    scale = np.linalg.norm(M) * np.linalg.norm(np.concatenate([va, vb]))
    assert np.max(np.abs(ya[s:] - ra)) <= 1e-13 * scale
    assert np.max(np.abs(yb[s:] - rb)) <= 1e-13 * scale


@pytest.mark.parametrize("n,s", [(130, 7), (700, 333), (1031, 65)])
def test_k1_matvec_evict_first_hint_is_bitwise_identical(n, s, monkeypatch):
    """the K1 instantiation whose matrix loads carry the L2 evict_first cache hint (ZQ_K1_EVICT) computes the same bits"""
    from tests import gpu_util as G
    M = O.gen_sym(n, 100 + n)
    rng = np.random.default_rng(n)
    va = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    vb = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    monkeypatch.delenv("ZQ_K1_EVICT", raising=False)
    y0 = G.matvec(M, s, va, vb)
    monkeypatch.setenv("ZQ_K1_EVICT", "1")
    y1 = G.matvec(M, s, va, vb)
    assert np.array_equal(y0[0][s:], y1[0][s:]) and np.array_equal(y0[1][s:], y1[1][s:])


def test_k1_ignores_upper_triangles():
    from tests import gpu_util as G
    n = 150
    M = O.gen_sym(n, 9)
    rng = np.random.default_rng(0)
    va = rng.standard_normal(n) + 0j
    vb = 1j * rng.standard_normal(n)
    y0 = G.matvec(M, 0, va, vb)
    M2 = M.copy()
    iu = np.triu_indices(n, 1)
    M2[:n, :n][iu] = np.nan
    M2[n:, :n][iu] = np.nan
    M2[n:, :n][np.diag_indices(n)] = np.nan
    y1 = G.matvec(M2, 0, va, vb)
    assert np.array_equal(y0[0], y1[0]) and np.array_equal(y0[1], y1[1])


@pytest.mark.parametrize("ta,tb,M,N,Kd,lower", [(0, 1, 100, 100, 128, 1), (0, 1, 67, 67, 40, 1), (1, 0, 64, 130, 300, 0),
                                                (0, 0, 200, 77, 64, 0), (0, 0, 5, 3, 2, 0), (1, 1, 33, 65, 17, 0),
                                                (0, 1, 129, 129, 256, 0)])
@pytest.mark.parametrize("three_m", [0, 1], ids=["4-product", "3M"])
def test_zgemm(ta, tb, M, N, Kd, lower, three_m):
    """both complex-product schemes, selected explicitly (the solver picks 3M for n >= 1024)"""
    from tests import gpu_util as G
    from zquatev_b200 import api
    api.lib().zq_test_set_gemm_3m(three_m)
    rng = np.random.default_rng(M * 7 + N)
    cr = lambda *s: rng.standard_normal(s) + 1j * rng.standard_normal(s)
    A = cr(Kd, M) if ta else cr(M, Kd)
    B = cr(N, Kd) if tb else cr(Kd, N)
    C = cr(M, N)
    alpha, beta = -1.0 + 0.5j, 1.0 - 0.25j
    got, _ = G.zgemm(ta, tb, alpha, A, B, beta, C, lower)
    opA = A.conj().T if ta else A
    opB = B.conj().T if tb else B
    ref = alpha * (opA @ opB) + beta * C
    if lower:
        ref = np.where(np.tril(np.ones((M, N), dtype=bool)), ref, C)
    assert np.max(np.abs(got - ref)) <= 1e-12 * Kd


@pytest.mark.parametrize("ta,tb,M,N,Kd,lower", [(0, 1, 100, 100, 128, 1), (0, 1, 67, 67, 40, 1), (1, 0, 64, 130, 300, 0),
                                                (0, 0, 200, 77, 64, 0), (0, 0, 5, 3, 2, 0), (1, 1, 33, 65, 17, 0),
                                                (0, 0, 129, 129, 7, 0), (1, 0, 1, 1, 1, 0)])
def test_qgemm8(ta, tb, M, N, Kd, lower):
    """K4/K6 quaternion GEMM with eight real products (qgemm.cu) against the 2 x 2 complex block form and against
    its numpy restatement (oracle.quat_kernels.qgemm8)."""
    from tests import gpu_util as G
    rng = np.random.default_rng(M * 11 + N + Kd)
    cr = lambda *s: rng.standard_normal(s) + 1j * rng.standard_normal(s)
    sa = (Kd, M) if ta else (M, Kd)
    sb = (N, Kd) if tb else (Kd, N)
    Aa, Ab, Ba, Bb, Ca, Cb = cr(*sa), cr(*sa), cr(*sb), cr(*sb), cr(M, N), cr(M, N)
    alpha, beta = -1.0, 1.0
    ga, gb, _ = G.qgemm(ta, tb, alpha, Aa, Ab, Ba, Bb, beta, Ca, Cb, lower)
    opA = (Aa.conj().T, -Ab.T) if ta else (Aa, Ab)          # quaternion conjugate transpose in pair form
    opB = (Ba.conj().T, -Bb.T) if tb else (Ba, Bb)
    ra, rb = K.qgemm_ref(opA[0], opA[1], opB[0], opB[1])
    ra, rb = alpha * ra + beta * Ca, alpha * rb + beta * Cb
    if lower:
        mask = np.tril(np.ones((M, N), dtype=bool))
        ra, rb = np.where(mask, ra, Ca), np.where(mask, rb, Cb)
    assert np.max(np.abs(ga - ra)) <= 2e-12 * Kd and np.max(np.abs(gb - rb)) <= 2e-12 * Kd
    fa, fb = K.qgemm8(opA[0], opA[1], opB[0], opB[1])       # same algorithm on the CPU: agreement to rounding of the sums
    fa, fb = alpha * fa + beta * Ca, alpha * fb + beta * Cb
    if lower:
        fa, fb = np.where(mask, fa, Ca), np.where(mask, fb, Cb)
    assert np.max(np.abs(ga - fa)) <= 2e-12 * Kd and np.max(np.abs(gb - fb)) <= 2e-12 * Kd


@pytest.mark.parametrize("tb,M,N,Kd,lower", [(1, 100, 100, 128, 1), (1, 67, 67, 40, 1), (0, 200, 77, 64, 0), (0, 5, 3, 2, 0),
                                             (0, 129, 129, 7, 0), (1, 33, 65, 17, 0), (0, 300, 260, 128, 0), (1, 257, 257, 100, 1),
                                             (0, 800, 790, 64, 0), (1, 801, 801, 128, 1)])   # > 444 tiles: several tiles per persistent CTA
def test_qgemm8_precombined_operands(tb, M, N, Kd, lower, monkeypatch):
    """the variant of the quaternion GEMM whose eight component sums are formed once per operand panel (qgemm8x.cu, ZQ_Q8X=1):
    against the 2 x 2 complex block form, and to rounding against the in-loop kernel behind the same door"""
    from tests import gpu_util as G
    rng = np.random.default_rng(M * 13 + N + Kd)
    cr = lambda *s: rng.standard_normal(s) + 1j * rng.standard_normal(s)
    sb = (N, Kd) if tb else (Kd, N)
    Aa, Ab, Ba, Bb, Ca, Cb = cr(M, Kd), cr(M, Kd), cr(*sb), cr(*sb), cr(M, N), cr(M, N)
    alpha, beta = -1.0, 1.0
    monkeypatch.setenv("ZQ_Q8X", "0")
    ya, yb, _ = G.qgemm(0, tb, alpha, Aa, Ab, Ba, Bb, beta, Ca, Cb, lower)
    monkeypatch.setenv("ZQ_Q8X", "1")
    ga, gb, _ = G.qgemm(0, tb, alpha, Aa, Ab, Ba, Bb, beta, Ca, Cb, lower)
    opB = (Ba.conj().T, -Bb.T) if tb else (Ba, Bb)
    ra, rb = K.qgemm_ref(Aa, Ab, opB[0], opB[1])
    ra, rb = alpha * ra + beta * Ca, alpha * rb + beta * Cb
    if lower:
        mask = np.tril(np.ones((M, N), dtype=bool))
        ra, rb = np.where(mask, ra, Ca), np.where(mask, rb, Cb)
    assert np.max(np.abs(ga - ra)) <= 2e-12 * Kd and np.max(np.abs(gb - rb)) <= 2e-12 * Kd
    assert np.max(np.abs(ga - ya)) <= 1e-13 * Kd and np.max(np.abs(gb - yb)) <= 1e-13 * Kd


@pytest.mark.parametrize("name,d,e", list(_tri_cases()))
def test_k8_stedc_cases(name, d, e):
    from tests import gpu_util as G
    rc, w, Z = G.stedc(d, e)
    assert rc == 0
    n = len(d)
    Tm = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
    wr = np.linalg.eigvalsh(Tm)
    nrm = max(np.abs(wr).max(), 1e-300)
    assert np.abs(w - wr).max() <= 50 * T.EPS * nrm
    assert np.linalg.norm(Tm @ Z - Z * w) / (n * nrm * T.EPS) < 2.0
    assert np.linalg.norm(Z.T @ Z - np.eye(n)) / (n * T.EPS) < 2.0


@pytest.mark.parametrize("n", [1, 2, 3, 31, 32, 33, 64, 65, 100, 513, 1000, 2049])
def test_k8_stedc_sizes(n):
    from tests import gpu_util as G
    rng = np.random.default_rng(n)
    d = rng.standard_normal(n)
    e = np.abs(rng.standard_normal(max(n - 1, 0)))
    rc, w, Z = G.stedc(d, e)
    assert rc == 0
    Tm = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
    wr = np.linalg.eigvalsh(Tm)
    nrm = np.abs(wr).max()
    assert np.abs(w - wr).max() <= 100 * T.EPS * nrm
    assert np.linalg.norm(Tm @ Z - Z * w) / (n * nrm * T.EPS) < 2.0
    assert np.linalg.norm(Z.T @ Z - np.eye(n)) / (n * T.EPS) < 2.0
    assert np.all(np.diff(w) >= 0)


@pytest.mark.parametrize("n", [1, 2, 50, 333, 1500])
def test_k9_bisect(n):
    from tests import gpu_util as G
    rng = np.random.default_rng(n)
    d = rng.standard_normal(n)
    e = rng.standard_normal(max(n - 1, 0))
    w = G.bisect(d, e)
    Tm = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
    wr = np.linalg.eigvalsh(Tm)
    assert np.abs(w - wr).max() <= 100 * T.EPS * max(np.abs(wr).max(), 1.0)   # << 1e-12 ||A||


@pytest.mark.parametrize("n,nb", [(2, 32), (3, 2), (21, 8), (22, 4), (64, 32), (65, 32), (130, 32), (200, 64), (300, 32)])
def test_tridiagonalisation_vs_oracle(n, nb, reduction_path):
    """K1-K4 chain == numpy restatement of the same formulation (different summation order only)."""
    from tests import gpu_util as G
    M = O.gen_sym(n, 200 + n)
    d, e, tau, al, Ah = G.tridiag(M, nb)
    dr, ala, alb, taur, Dr, Er = K.tridiagonalise(M[:n, :n], M[n:, :n], nb)
    er = np.sqrt(np.abs(ala) ** 2 + np.abs(alb) ** 2)
    nrm = np.linalg.norm(M, 2)
    assert np.max(np.abs(d - dr)) <= 1e-12 * nrm
    assert np.max(np.abs(e[: n - 1] - er)) <= 1e-12 * nrm
    assert np.max(np.abs(tau[: n - 1] - taur)) <= 1e-10
    # eigenvalues of the real tridiagonal == eigenvalues of M (each twice)
    Tm = np.diag(d) + np.diag(e[: n - 1], 1) + np.diag(e[: n - 1], -1)
    wm = np.linalg.eigvalsh(M)[0::2]
    assert np.max(np.abs(np.linalg.eigvalsh(Tm) - wm)) <= 1e-12 * nrm
    # reflector tails left in the lower triangles
    low = np.tril_indices(n, -2)
    if len(low[0]):
        assert np.max(np.abs(Ah[:n][low] - Dr[low])) <= 1e-10
        assert np.max(np.abs(Ah[n:][low] - Er[low])) <= 1e-10
