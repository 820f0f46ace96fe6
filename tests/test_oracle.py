"""CPU tests (-m "not gpu"): the oracle is pinned against (a) the golden vectors generated from
the unmodified reference (tests/golden) and (b), when oracle/_ref has been
built in this container, the reference itself."""
import json
import os

import numpy as np
import pytest

from oracle import quat_kernels as K
from oracle import tridiag_dc as T
from oracle import zquatev_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")
TESTCC = json.load(open(os.path.join(GOLD, "testcc_eigs.json")))
SYM = json.load(open(os.path.join(GOLD, "sym_eigs.json")))

# SURVEY.md Appendix B (12 digits), independent of the JSON fixtures
APPENDIX_B = {1: (0.1178, None, 0.1178), 2: (-0.842613071032, 1.549713071032, 1.549713071032),
              3: (-1.560137575611, 0.044679547778, 2.799058027833),
              21: (-9.475257015121, -5.004923945245, 17.984680164637),
              22: (-9.901600859158, -5.101077367863, 18.751733289038),
              23: (-10.350995860151, -5.218486936616, 19.481924790484),
              64: (-26.643093218738, -12.668480878225, 53.670043618202),
              200: (-83.066705628507, -35.482307890128, 165.565405199418),
              500: (-207.859872474066, -84.647193346768, 414.425661615841)}


def test_generator_first_element():
    A, B, C = O.gen_testcc(3)
    assert abs(A[0, 0] - 0.1178) < 1e-15
    assert np.allclose(C, C.conj().T)
    assert np.allclose(B, -B.T)


@pytest.mark.parametrize("n", [1, 2, 3, 21, 22, 23, 64, 200, 500])
def test_golden_matches_appendix_b(n):
    eig = np.array(TESTCC[str(n)]["eig"])
    e0, e1, el = APPENDIX_B[n]
    assert abs(eig[0] - e0) < 1e-11 and abs(eig[-1] - el) < 1e-11
    if e1 is not None:
        assert abs(eig[1] - e1) < 1e-11
    A, _, _ = O.gen_testcc(n)
    assert abs(eig.sum() - A.trace().real) < 1e-10 * max(1.0, n)


@pytest.mark.parametrize("n", [1, 2, 3, 21, 22, 23, 64])
def test_restatement_vs_golden(n):
    """numpy restatement of unblocked.cc/zquatev.cc reproduces the reference's eigenvalues."""
    _, _, C = O.gen_testcc(n)
    eig, out, info = O.ref_zquatev(C)
    g = TESTCC[str(n)]
    assert info == 0
    assert np.max(np.abs(eig - np.array(g["eig"]))) <= 1e-12 * g["two_norm"]
    res, orth, pair = O.quality(C, out, eig)
    assert pair == 0.0 and res < 2.0 and orth < 5.0


@pytest.mark.parametrize("key", ["5_32", "33_33", "100_34"])
def test_restatement_vs_golden_sym(key):
    n, seed = (int(x) for x in key.split("_"))
    M = O.gen_sym(n, seed)
    eig, out, _ = O.ref_zquatev(M)
    g = SYM[key]
    assert np.max(np.abs(eig - np.array(g["eig"]))) <= 1e-12 * g["two_norm"]


@pytest.mark.parametrize("n,nb", [(1, 4), (2, 4), (3, 2), (21, 8), (22, 4), (23, 32), (64, 8), (64, 32), (200, 32)])
def test_b200_formulation_vs_golden(n, nb):
    """kernel-level oracle (quaternion reflectors, lower triangles, compact WY) == reference."""
    _, _, C = O.gen_testcc(n)
    eig, out = K.solve(C, nb)
    g = TESTCC[str(n)]
    assert np.max(np.abs(eig - np.array(g["eig"]))) <= 1e-12 * g["two_norm"]
    res, orth, pair = O.quality(C, out, eig)
    assert pair == 0.0
    assert res <= max(2 * g["residual"], 0.5) and orth <= max(2 * g["orthogonality"], 2.0)


def test_matvec_lower_equals_full():
    M = O.gen_sym(37, 5)
    D, E = M[:37, :37], M[37:, :37]
    rng = np.random.default_rng(1)
    va = rng.standard_normal(37) + 1j * rng.standard_normal(37)
    vb = rng.standard_normal(37) + 1j * rng.standard_normal(37)
    ya, yb = K.matvec_full(D, E, va, vb)
    za, zb = K.matvec_lower(D, E, va, vb)
    assert np.allclose(ya, za, atol=1e-13) and np.allclose(yb, zb, atol=1e-13)
    y = M @ np.concatenate([va, vb])
    assert np.allclose(np.concatenate([ya, yb]), y, atol=1e-13)


@pytest.mark.skipif(not O.RefLib.available(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("n", [2, 23, 40, 77])
def test_restatement_vs_reference_lib(n):
    ref = O.RefLib()
    M = O.gen_sym(n, 7)
    e1, o1, i1 = ref.zquatev(M)
    e2, o2, i2 = O.ref_zquatev(M)
    e3, o3 = K.solve(M, 8)
    nrm = np.max(np.abs(e1))
    assert i1 == 0 and i2 == 0
    assert np.max(np.abs(e1 - e2)) <= 1e-12 * nrm and np.max(np.abs(e1 - e3)) <= 1e-12 * nrm
    for o, e in ((o1, e1), (o2, e2), (o3, e3)):
        res, orth, pair = O.quality(M, o, e)
        assert pair == 0.0 and res < 2.0 and orth < 5.0


@pytest.mark.skipif(not O.RefLib.available(), reason="oracle/_ref not built")
def test_reference_lib_matches_golden():
    ref = O.RefLib()
    _, _, C = O.gen_testcc(64)
    eig, out, info = ref.zquatev(C)
    assert info == 0
    assert np.max(np.abs(eig - np.array(TESTCC["64"]["eig"]))) < 1e-11


def _tri_cases():
    rng = np.random.default_rng(0)
    yield "random", rng.standard_normal(150), np.abs(rng.standard_normal(149))
    n = 101
    yield "wilkinson", np.abs(np.arange(n) - n // 2).astype(float), np.ones(n - 1)
    yield "identity", np.ones(70), np.zeros(69)
    yield "clusters", np.repeat([1.0, 2.0, 3.0, 4.0], 25), 1e-9 * np.abs(rng.standard_normal(99))
    w21 = np.abs(np.arange(21) - 10).astype(float)
    e = np.ones(104)
    e[20::21] = 1e-10
    yield "glued", np.tile(w21, 5), e
    yield "neg_e", rng.standard_normal(64), rng.standard_normal(63)


@pytest.mark.parametrize("name,d,e", list(_tri_cases()))
def test_dc_prototype_vs_lapack(name, d, e):
    w, Z = T.stedc(d, e, leaf=8)
    n = len(d)
    Tm = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
    wr = np.linalg.eigvalsh(Tm)
    nrm = max(np.abs(wr).max(), 1e-300)
    assert np.abs(w - wr).max() <= 50 * T.EPS * nrm
    assert np.linalg.norm(Tm @ Z - Z * w) / (n * nrm * T.EPS) < 2.0
    assert np.linalg.norm(Z.T @ Z - np.eye(n)) / (n * T.EPS) < 2.0


@pytest.mark.parametrize("n,nb", [(1, 4), (2, 4), (3, 2), (7, 3), (40, 8), (65, 64), (100, 32)])
def test_one_cta_restatement_vs_blocked(n, nb):
    """K5's data flow (unblocked, rank-2 update fused with the next mat-vec, Gram columns read back from the
    reflector tails) reaches the same tridiagonal, reflectors and T-factor inputs as the blocked restatement."""
    M = O.gen_sym(n, 7)
    D, E = M[:n, :n], M[n:, :n]
    d0, aa0, ab0, t0, D0, E0 = K.tridiagonalise(D, E, nb)
    d1, aa1, ab1, t1, D1, E1, Ga, Gb = K.tridiagonalise_one_cta(D, E, nb)
    nrm = max(np.linalg.norm(M, 2), 1.0)
    assert np.max(np.abs(d0 - d1)) <= 1e-12 * nrm
    if n == 1:
        return
    e0 = np.sqrt(np.abs(aa0) ** 2 + np.abs(ab0) ** 2)
    e1 = np.sqrt(np.abs(aa1) ** 2 + np.abs(ab1) ** 2)
    assert np.max(np.abs(e0 - e1)) <= 1e-12 * nrm and np.max(np.abs(t0 - t1)) <= 1e-10
    assert np.max(np.abs(np.tril(D0, -2) - np.tril(D1, -2))) <= 1e-10
    assert np.max(np.abs(np.tril(E0, -2) - np.tril(E1, -2))) <= 1e-10
    for j0 in range(0, n - 1, nb):                       # G[k, t] == (V^H V)[t, i] of the panel
        kb = min(nb, n - 1 - j0)
        P = K.phi_panel(D1, E1, j0, kb)
        m = n - 1 - j0
        Va, Vb = P[:m, :kb], P[m:, :kb]
        for i in range(kb):
            for t in range(i):
                ga, gb = K.PH(Va[:, t:t + 1], Vb[:, t:t + 1], Va[:, i], Vb[:, i])
                assert abs(ga[0] - Ga[j0 + i, t]) <= 1e-13 and abs(gb[0] - Gb[j0 + i, t]) <= 1e-13


@pytest.mark.parametrize("n,nb", [(2, 4), (5, 2), (9, 4), (33, 8), (70, 16), (64, 64), (130, 32)])
def test_paired_backtransform_restatement(n, nb):
    """two panels merged per step (T12 = [[Ta, -Ta Pa^H Pb Tb], [0, Tb]]) == panel-by-panel back-transformation"""
    rng = np.random.default_rng(n)
    M = O.gen_sym(n, 3)
    _, _, _, tau, Df, Ef = K.tridiagonalise(M[:n, :n], M[n:, :n], nb)
    Xa = rng.normal(size=(n, n)) + 1j * rng.normal(size=(n, n))
    Xb = rng.normal(size=(n, n)) + 1j * rng.normal(size=(n, n))
    a0, b0 = K.backtransform(Df, Ef, tau, Xa.copy(), Xb.copy(), nb)
    a1, b1 = K.backtransform_paired(Df, Ef, tau, Xa.copy(), Xb.copy(), nb)
    scale = max(np.abs(a0).max(), np.abs(b0).max())
    assert np.abs(a0 - a1).max() <= 1e-13 * scale and np.abs(b0 - b1).max() <= 1e-13 * scale


def test_qgemm8_restatement_and_backtransform():
    """the eight-product quaternion GEMM (csrc/qgemm.cu restated in oracle.quat_kernels.qgemm8) equals the 2 x 2 complex
    block product, and the back-transformation built on it keeps the quality of the stacked complex form"""
    rng = np.random.default_rng(5)
    cr = lambda *s: rng.standard_normal(s) + 1j * rng.standard_normal(s)
    Aa, Ab, Ba, Bb = cr(13, 9), cr(13, 9), cr(9, 11), cr(9, 11)
    ra, rb = K.qgemm_ref(Aa, Ab, Ba, Bb)
    fa, fb = K.qgemm8(Aa, Ab, Ba, Bb)
    assert np.max(np.abs(ra - fa)) < 1e-13 and np.max(np.abs(rb - fb)) < 1e-13
    n, nb = 150, 16
    M = O.gen_sym(n, 9)
    d, ala, alb, tau, Df, Ef = K.tridiagonalise(np.array(M[:n, :n]), np.array(M[n:, :n]), nb)
    e, sa, sb = K.phase_chain(ala, alb)
    w, Z = np.linalg.eigh(np.diag(d) + np.diag(e, -1) + np.diag(e, 1))
    res = {}
    for name, bt in (("stacked", K.backtransform), ("q8", K.backtransform_q8)):
        Xa, Xb = bt(Df, Ef, tau, sa[:, None] * Z, sb[:, None] * Z, nb)
        out = np.empty((2 * n, 2 * n), dtype=np.complex128)
        out[:n, :n], out[n:, :n] = Xa, Xb
        out[:n, n:], out[n:, n:] = -np.conj(Xb), np.conj(Xa)
        res[name] = O.quality(M, out, w)
    assert res["q8"][2] == 0.0
    assert res["q8"][0] <= 1.5 * res["stacked"][0] + 0.01 and res["q8"][1] <= 1.5 * res["stacked"][1] + 0.05, res


def test_qgemm8_precombined_planes_restatement():
    """the pre-combined-operand form of the eight-product GEMM (csrc/qgemm8x.cu: component sums once per operand into zero-padded
    real planes, eight plain real products, recombination) equals the in-loop form and the 2 x 2 complex block product --
    also for the conjugate-transposed right operand of the trailing update and for ragged sizes (padding)"""
    rng = np.random.default_rng(8)
    cr = lambda *s: rng.standard_normal(s) + 1j * rng.standard_normal(s)
    for (M, N, Kd) in [(13, 11, 9), (32, 64, 8), (67, 67, 40), (5, 3, 2)]:
        Aa, Ab, Ba, Bb = cr(M, Kd), cr(M, Kd), cr(Kd, N), cr(Kd, N)
        A8, B8 = K.q8_planes_a(Aa, Ab), K.q8_planes_b(Ba, Bb)
        assert A8.shape[1] % 8 == 0 and A8.shape[2] % 32 == 0 and B8.shape[2] % 32 == 0
        ga, gb = K.qgemm8_planes(A8, B8, M, N)
        ra, rb = K.qgemm_ref(Aa, Ab, Ba, Bb)
        fa, fb = K.qgemm8(Aa, Ab, Ba, Bb)
        assert np.max(np.abs(ga - ra)) < 1e-12 and np.max(np.abs(gb - rb)) < 1e-12
        assert np.max(np.abs(ga - fa)) < 1e-13 and np.max(np.abs(gb - fb)) < 1e-13
        # trailing update: right operand = quaternion conjugate transpose of a stored N x K array S
        Sa, Sb = cr(N, Kd), cr(N, Kd)
        opa, opb = np.conj(Sa).T, -Sb.T
        ga, gb = K.qgemm8_planes(A8, K.q8_planes_b(opa, opb), M, N)
        ra, rb = K.qgemm_ref(Aa, Ab, opa, opb)
        assert np.max(np.abs(ga - ra)) < 1e-12 and np.max(np.abs(gb - rb)) < 1e-12
