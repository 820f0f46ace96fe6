"""Kernel-level CPU oracle for the B200 formulation.  TEST INFRASTRUCTURE ONLY.

The CUDA path does not replay the reference's H1 / Givens / H2 column step
(unblocked.cc:53-128); it uses ONE quaternion Householder reflector per column, which
reaches the same tridiagonal form Q^H M Q = diag(T, T*) that blocked.cc:42-68 targets.
This module restates every device kernel of that formulation in numpy so the
`-m gpu` tests can compare kernel by kernel (K1 mat-vec, panel, K4 trailing update,
phase chain, T factor, K6 back-transform, K10 pairing).  The whole chain is itself
checked against the reference (oracle/_ref) in tests/test_oracle.py.

Conventions: a quaternion q = a + j b is the complex pair (a, b);
matrices act from the left, scalars multiply vectors from the right;
Phi(D,E) = [[D, -conj(E)], [E, conj(D)]].
"""
from __future__ import annotations

import numpy as np

c = np.conj


def qmul(p, q):
    (a, b), (cc, d) = p, q
    return a * cc - c(b) * d, b * cc + c(a) * d


def qconj(p):
    a, b = p
    return c(a), -b


def matvec_full(D, E, va, vb):
    """K1 semantics on full storage: y = (D + jE)(va + j vb)."""
    return D @ va - c(E) @ vb, E @ va + c(D) @ vb


def matvec_lower(D, E, va, vb):
    """K1 as the device kernel computes it: only the LOWER triangles of D (Hermitian,
    real diagonal) and E (antisymmetric, zero diagonal) are read."""
    Dl = np.tril(D, -1)
    El = np.tril(E, -1)
    dg = D.diagonal().real
    ya = Dl @ va - c(El) @ vb + c(Dl).T @ va + c(El).T @ vb + dg * va
    yb = El @ va + c(Dl) @ vb - El.T @ va + Dl.T @ vb + dg * vb
    return ya, yb


def PH(Xa, Xb, ya, yb):
    """X^H y for a quaternion panel X and quaternion vector y."""
    return c(Xa).T @ ya + c(Xb).T @ yb, Xa.T @ yb - Xb.T @ ya


def PV(Xa, Xb, ca, cb):
    """X c for a quaternion panel X and quaternion coefficient vector c."""
    return Xa @ ca - c(Xb) @ cb, Xb @ ca + c(Xa) @ cb


def make_reflector(xa, xb):
    """Quaternion Householder: H = I - tau v v^H (tau REAL, v[0] = 1) with H x = e1 alpha.
    Returns (alpha_a, alpha_b, tau, va, vb)."""
    nx2 = float(np.sum(xa.real ** 2 + xa.imag ** 2 + xb.real ** 2 + xb.imag ** 2))
    x1 = np.sqrt(abs(xa[0]) ** 2 + abs(xb[0]) ** 2)
    rest2 = nx2 - x1 * x1
    va, vb = xa.copy(), xb.copy()
    if nx2 == 0.0 or (rest2 <= 0.0 and len(xa) == 1 and False):
        va[:] = 0
        vb[:] = 0
        va[0] = 1.0
        return 0j, 0j, 0.0, va, vb
    nx = np.sqrt(nx2)
    if x1 == 0.0:
        pa, pb = 1.0 + 0j, 0j
    else:
        pa, pb = xa[0] / x1, xb[0] / x1
    al_a, al_b = -pa * nx, -pb * nx
    u1a, u1b = xa[0] - al_a, xb[0] - al_b
    u1n2 = abs(u1a) ** 2 + abs(u1b) ** 2
    # ||u||^2 = ||x||^2 + 2 |x1| ||x|| + ||x||^2 - ... computed directly
    un2 = nx2 - x1 * x1 + u1n2
    tau = 2.0 * u1n2 / un2
    ia, ib = c(u1a) / u1n2, -u1b / u1n2          # u1^{-1}
    va, vb = qmul((xa, xb), (ia, ib))              # right scale
    va[0], vb[0] = 1.0, 0.0
    return al_a, al_b, tau, va, vb


def tridiagonalise(D, E, nb):
    """Blocked reduction (panel + stacked rank-2k update), lower triangles only.

    Returns d[n], alpha_a[n-1], alpha_b[n-1], tau[n-1] and the arrays D, E whose strictly
    sub-sub-diagonal parts now hold the reflector tails v[1:] (v[0] = 1 implicit)."""
    D = D.copy()
    E = E.copy()
    n = D.shape[0]
    d = np.zeros(n)
    ala = np.zeros(max(n - 1, 0), dtype=np.complex128)
    alb = np.zeros(max(n - 1, 0), dtype=np.complex128)
    tau = np.zeros(max(n - 1, 0))
    for j0 in range(0, n - 1, nb):
        kb = min(nb, n - 1 - j0)
        Va = np.zeros((n, kb), dtype=np.complex128)
        Vb = np.zeros_like(Va)
        Wa = np.zeros_like(Va)
        Wb = np.zeros_like(Va)
        for i in range(kb):
            k = j0 + i
            ca, cb = D[k:, k].copy(), E[k:, k].copy()
            if i > 0:
                qa, qb = qconj((Wa[k, :i], Wb[k, :i]))
                ta, tb = PV(Va[k:, :i], Vb[k:, :i], qa, qb)
                ca -= ta
                cb -= tb
                qa, qb = qconj((Va[k, :i], Vb[k, :i]))
                ta, tb = PV(Wa[k:, :i], Wb[k:, :i], qa, qb)
                ca -= ta
                cb -= tb
            d[k] = ca[0].real
            al_a, al_b, t, va, vb = make_reflector(ca[1:], cb[1:])
            ala[k], alb[k], tau[k] = al_a, al_b, t
            D[k + 1:, k] = va
            E[k + 1:, k] = vb
            fa = np.zeros(n, dtype=np.complex128)
            fb = np.zeros(n, dtype=np.complex128)
            fa[k + 1:], fb[k + 1:] = va, vb
            # mat-vec with the matrix as of the START of the panel, trailing block only
            pa = np.zeros(n, dtype=np.complex128)
            pb = np.zeros(n, dtype=np.complex128)
            pa[k + 1:], pb[k + 1:] = matvec_lower(D[k + 1:, k + 1:], E[k + 1:, k + 1:], va, vb)
            if i > 0:
                ga, gb = PH(Wa[:, :i], Wb[:, :i], fa, fb)
                ta, tb = PV(Va[:, :i], Vb[:, :i], ga, gb)
                pa -= ta
                pb -= tb
                ga, gb = PH(Va[:, :i], Vb[:, :i], fa, fb)
                ta, tb = PV(Wa[:, :i], Wb[:, :i], ga, gb)
                pa -= ta
                pb -= tb
            pa *= t
            pb *= t
            pa[:k + 1] = 0
            pb[:k + 1] = 0
            g = float((np.vdot(fa, pa) + np.vdot(fb, pb)).real)
            Va[:, i], Vb[:, i] = fa, fb
            Wa[:, i] = pa - 0.5 * t * g * fa
            Wb[:, i] = pb - 0.5 * t * g * fb
        r0 = j0 + kb
        if r0 < n:
            L = np.block([[Va[r0:], c(Vb[r0:]), Wa[r0:], c(Wb[r0:])],
                          [Vb[r0:], -c(Va[r0:]), Wb[r0:], -c(Wa[r0:])]])
            R = np.hstack([Wa[r0:], c(Wb[r0:]), Va[r0:], c(Vb[r0:])])
            U = L @ c(R).T
            m = n - r0
            D[r0:, r0:] -= np.tril(U[:m])
            E[r0:, r0:] -= np.tril(U[m:], -1)
    d[n - 1] = D[n - 1, n - 1].real
    return d, ala, alb, tau, D, E


def tridiagonalise_one_cta(D, E, nb):
    """K5 (small.cu) restated step by step: unblocked reduction of ONE small matrix in which the pass
    that applies the rank-2 update of column k-1 (M -= v_{k-1} w_{k-1}^H + w_{k-1} v_{k-1}^H, lower
    triangles) also forms y = M v_k from the updated entries.  Same reflector formulas as
    `tridiagonalise`, so d / alpha / tau / reflector tails agree to rounding for every nb.
    Additionally returns G[k, t] = V_t^H v_k (t < k - j0: earlier reflectors of the same panel),
    the quantity the device keeps for the compact-WY T factors."""
    D = D.copy()
    E = E.copy()
    n = D.shape[0]
    d = np.zeros(n)
    ala = np.zeros(max(n - 1, 0), dtype=np.complex128)
    alb = np.zeros(max(n - 1, 0), dtype=np.complex128)
    tau = np.zeros(max(n - 1, 0))
    Ga = np.zeros((n, nb), dtype=np.complex128)
    Gb = np.zeros((n, nb), dtype=np.complex128)
    vpa = np.zeros(n, dtype=np.complex128)
    vpb = np.zeros(n, dtype=np.complex128)
    wa = np.zeros(n, dtype=np.complex128)
    wb = np.zeros(n, dtype=np.complex128)

    def rank2(ra, rb, cidx):
        """(vp wv^H + wv vp^H)[rows, cidx] as a quaternion (a, b) column"""
        q1 = qmul((vpa[ra:rb], vpb[ra:rb]), qconj((wa[cidx], wb[cidx])))
        q2 = qmul((wa[ra:rb], wb[ra:rb]), qconj((vpa[cidx], vpb[cidx])))
        return q1[0] + q2[0], q1[1] + q2[1]

    for k in range(n - 1):
        s = k + 1
        j0 = (k // nb) * nb
        i = k - j0
        # (a) column k, fully updated
        ca, cb = D[k:, k].copy(), E[k:, k].copy()
        if k > 0:
            ua, ub = rank2(k, n, k)
            ca -= ua
            cb -= ub
        d[k] = ca[0].real
        # (b) reflector
        al_a, al_b, t, va_, vb_ = make_reflector(ca[1:], cb[1:])
        ala[k], alb[k], tau[k] = al_a, al_b, t
        D[s:, k] = va_
        E[s:, k] = vb_
        va = np.zeros(n, dtype=np.complex128)
        vb = np.zeros(n, dtype=np.complex128)
        va[s:], vb[s:] = va_, vb_
        # (b2) Gram column: tails of the panel's earlier reflectors are read back from D, E (rows >= s)
        for tt in range(i):
            cj = j0 + tt
            ga, gb = PH(D[s:, cj:cj + 1], E[s:, cj:cj + 1], va[s:], vb[s:])
            Ga[k, tt], Gb[k, tt] = ga[0], gb[0]
        # (c) fused pass over the lower triangles of rows/cols [s, n)
        if k > 0:
            for cc in range(s, n):
                ua, ub = rank2(cc, n, cc)
                D[cc:, cc] -= ua
                E[cc:, cc] -= ub
                D[cc, cc] = D[cc, cc].real
                E[cc, cc] = 0.0
        ya, yb = matvec_lower(D[s:, s:], E[s:, s:], va[s:], vb[s:])
        # (d) p, g, w
        pa = np.zeros(n, dtype=np.complex128)
        pb = np.zeros(n, dtype=np.complex128)
        pa[s:], pb[s:] = t * ya, t * yb
        g = float((np.vdot(va, pa) + np.vdot(vb, pb)).real)
        wa = pa - 0.5 * t * g * va
        wb = pb - 0.5 * t * g * vb
        vpa, vpb = va, vb
    last = D[n - 1, n - 1]
    if n > 1:
        ua, ub = rank2(n - 1, n, n - 1)
        last = last - ua[0]
    d[n - 1] = last.real
    # rows/cols of the tails: leave only what the device leaves meaningful (strictly sub-sub-diagonal)
    return d, ala, alb, tau, D, E, Ga, Gb


def phase_chain(ala, alb):
    """s_0 = 1, s_{k+1} = (alpha_k / |alpha_k|) s_k -- the diagonal unit-quaternion
    similarity that makes the quaternion tridiagonal real symmetric (e_k = |alpha_k|)."""
    n = len(ala) + 1
    sa = np.zeros(n, dtype=np.complex128)
    sb = np.zeros(n, dtype=np.complex128)
    sa[0] = 1.0
    e = np.sqrt(np.abs(ala) ** 2 + np.abs(alb) ** 2)
    for k in range(n - 1):
        if e[k] == 0.0:
            sa[k + 1], sb[k + 1] = sa[k], sb[k]
        else:
            sa[k + 1], sb[k + 1] = qmul((ala[k] / e[k], alb[k] / e[k]), (sa[k], sb[k]))
            nrm = np.sqrt(abs(sa[k + 1]) ** 2 + abs(sb[k + 1]) ** 2)
            sa[k + 1] /= nrm
            sb[k + 1] /= nrm
    return e, sa, sb


def phi_panel(D, E, j0, kb):
    """Phi(V) for the reflectors of panel [j0, j0+kb): complex (2m x 2kb), rows j0+1..n-1,
    m = n-1-j0; columns [V | Theta V] = [[Va, -conj(Vb)], [Vb, conj(Va)]]."""
    n = D.shape[0]
    m = n - 1 - j0
    Va = np.zeros((m, kb), dtype=np.complex128)
    Vb = np.zeros((m, kb), dtype=np.complex128)
    for i in range(kb):
        Va[i, i] = 1.0
        Va[i + 1:, i] = D[j0 + i + 2:, j0 + i]
        Vb[i + 1:, i] = E[j0 + i + 2:, j0 + i]
    return np.block([[Va, -c(Vb)], [Vb, c(Va)]])


def tfactor(P, tau):
    """Compact-WY T (2kb x 2kb complex, the Phi-image of the upper-triangular quaternion T)
    with H_0 H_1 ... H_{kb-1} = I - P T P^H, column order of P = [V | Theta V]."""
    kb = len(tau)
    G = c(P).T @ P
    T = np.zeros((2 * kb, 2 * kb), dtype=np.complex128)
    for i in range(kb):
        idx_prev = np.concatenate([np.arange(i), kb + np.arange(i)]).astype(int)
        for col in (i, kb + i):
            T[col, col] = tau[i]
            if i > 0:
                T[np.ix_(idx_prev, [col])] = -tau[i] * (T[np.ix_(idx_prev, idx_prev)] @ G[np.ix_(idx_prev, [col])])
    return T


def backtransform(D, E, tau, Xa, Xb, nb):
    """K6: X <- H_0 H_1 ... H_{n-2} X, panel by panel in reverse order, as stacked complex
    GEMMs on [Xa; Xb] with Phi(V)."""
    n = D.shape[0]
    X = np.vstack([Xa, Xb])
    starts = list(range(0, n - 1, nb))
    for j0 in reversed(starts):
        kb = min(nb, n - 1 - j0)
        m = n - 1 - j0
        P = phi_panel(D, E, j0, kb)
        T = tfactor(P, tau[j0:j0 + kb])
        rows = np.concatenate([np.arange(j0 + 1, n), n + np.arange(j0 + 1, n)])
        Y = c(P).T @ X[rows]
        X[rows] -= P @ (T @ Y)
    return X[:n], X[n:]


def backtransform_paired(D, E, tau, Xa, Xb, nb):
    """K6 with two panels merged per step (solver.cu `backtransform`, ZQ_BT_PAIR=1): for panels j < j' applied
    together,  H_j H_j' = I - [P_j P_j'] [[T_j, -T_j (P_j^H P_j') T_j'], [0, T_j']] [P_j P_j']^H,
    where P_j' is zero-padded to the rows of P_j.  The K = 4 nb update GEMM visits every C tile half as often."""
    n = D.shape[0]
    X = np.vstack([Xa, Xb])
    starts = list(range(0, n - 1, nb))
    i = len(starts) - 1
    while i >= 0:
        if i >= 1:
            ja, jb = starts[i - 1], starts[i]
            ka, kbb = min(nb, n - 1 - ja), min(nb, n - 1 - jb)
            ma, mb = n - 1 - ja, n - 1 - jb
            Pa = phi_panel(D, E, ja, ka)
            Pb0 = phi_panel(D, E, jb, kbb)
            Pb = np.zeros((2 * ma, 2 * kbb), dtype=np.complex128)
            Pb[ma - mb:ma] = Pb0[:mb]                # a-rows, shifted down by jb - ja
            Pb[2 * ma - mb:] = Pb0[mb:]              # b-rows
            Ta = tfactor(Pa, tau[ja:ja + ka])
            Tb = tfactor(Pb0, tau[jb:jb + kbb])
            S = c(Pa).T @ Pb
            Pc = np.hstack([Pa, Pb])
            T12 = np.zeros((2 * (ka + kbb), 2 * (ka + kbb)), dtype=np.complex128)
            T12[:2 * ka, :2 * ka] = Ta
            T12[2 * ka:, 2 * ka:] = Tb
            T12[:2 * ka, 2 * ka:] = -Ta @ (S @ Tb)
            rows = np.concatenate([np.arange(ja + 1, n), n + np.arange(ja + 1, n)])
            Y = c(Pc).T @ X[rows]
            X[rows] -= Pc @ (T12 @ Y)
            i -= 2
        else:
            j0 = starts[0]
            kb = min(nb, n - 1 - j0)
            P = phi_panel(D, E, j0, kb)
            T = tfactor(P, tau[j0:j0 + kb])
            rows = np.concatenate([np.arange(j0 + 1, n), n + np.arange(j0 + 1, n)])
            Y = c(P).T @ X[rows]
            X[rows] -= P @ (T @ Y)
            i -= 1
    return X[:n], X[n:]


def solve(M, nb=8, tridiag_solver=None):
    """Whole B200 formulation on the CPU: returns (eig, out) like ts::zquatev."""
    n = M.shape[0] // 2
    D = np.array(M[:n, :n])
    E = np.array(M[n:, :n])
    d, ala, alb, tau, Df, Ef = tridiagonalise(D, E, nb)
    e, sa, sb = phase_chain(ala, alb)
    if tridiag_solver is None:
        T = np.diag(d) + np.diag(e, -1) + np.diag(e, 1)
        w, Z = np.linalg.eigh(T)
    else:
        w, Z = tridiag_solver(d, e)
    Xa, Xb = sa[:, None] * Z, sb[:, None] * Z
    Xa, Xb = backtransform(Df, Ef, tau, Xa, Xb, nb)
    out = np.empty((2 * n, 2 * n), dtype=np.complex128)
    out[:n, :n], out[n:, :n] = Xa, Xb
    out[:n, n:], out[n:, n:] = -c(Xb), c(Xa)
    return w, out


# --------------------------------------------------------------------------------------
# quaternion GEMM with eight real products (csrc/qgemm.cu restated)
# --------------------------------------------------------------------------------------
def q_components(Qa, Qb):
    """(a, b) complex pair of a quaternion matrix Q = Qa + j Qb -> its four real component planes (1, i, j, k)."""
    return [Qa.real.copy(), Qa.imag.copy(), Qb.real.copy(), -Qb.imag.copy()]


def q_from_components(q):
    return q[0] + 1j * q[1], q[2] - 1j * q[3]


def q_conj_transpose(q):
    """component planes of Q^H (quaternion conjugate transpose)"""
    return [q[0].T, -q[1].T, -q[2].T, -q[3].T]


def qgemm8(Aa, Ab, Ba, Bb):
    """C = A B for quaternion matrices in complex-pair form with EIGHT real matrix products (bilinear rank of the
    quaternion algebra; the products keep the a-combination on the left, so the identity holds for matrices):
    exactly the combinations of csrc/qgemm.cu (combos_a / combos_b / epilogue).  Returns (Ca, Cb)."""
    a1, a2, a3, a4 = q_components(Aa, Ab)
    b1, b2, b3, b4 = q_components(Ba, Bb)
    p1 = (a4 + a2) @ (b2 + b3)
    p2 = (a1 - a3) @ (b1 + b4)
    p3 = (a1 + a3) @ (b1 - b4)
    p4 = (a4 - a2) @ (b2 - b3)
    p5 = (a4 - a3) @ (b3 - b4)
    p6 = (a2 + a1) @ (b2 + b1)
    p7 = (a1 - a2) @ (b3 + b4)
    p8 = (a4 + a3) @ (b1 - b2)
    s123 = (p1 + p2) + p3
    s = 0.5 * (s123 + p4)
    return q_from_components([(s - p1) + p5, (s - s123) + p6, (s - p2) + p7, (s - p3) + p8])


def q8_planes_a(Aa, Ab, row_pad=32, k_pad=8):
    """csrc/qgemm8x.cu k_combine_a: the eight LEFT component sums of A (M x K quaternions) as real planes
    A8[e][k][m], rows zero-padded to a multiple of 32 and k to a multiple of 8 (combos_a of gemm_tiles.cuh)."""
    a1, a2, a3, a4 = q_components(Aa, Ab)
    sums = [a4 + a2, a1 - a3, a1 + a3, a4 - a2, a4 - a3, a2 + a1, a1 - a2, a4 + a3]
    M, Kd = Aa.shape
    ld, k8 = -(-M // row_pad) * row_pad, -(-Kd // k_pad) * k_pad
    A8 = np.zeros((8, k8, ld))
    for e, pl in enumerate(sums):
        A8[e, :Kd, :M] = pl.T
    return A8


def q8_planes_b(Ba, Bb, col_pad=32, k_pad=8):
    """csrc/qgemm8x.cu k_combine_b: the eight RIGHT component sums of B (K x N quaternions) as real planes B8[e][k][n]."""
    b1, b2, b3, b4 = q_components(Ba, Bb)
    sums = [b2 + b3, b1 + b4, b1 - b4, b2 - b3, b3 - b4, b2 + b1, b3 + b4, b1 - b2]
    Kd, N = Ba.shape
    ld, k8 = -(-N // col_pad) * col_pad, -(-Kd // k_pad) * k_pad
    B8 = np.zeros((8, k8, ld))
    for e, pl in enumerate(sums):
        B8[e, :Kd, :N] = pl
    return B8


def qgemm8_planes(A8, B8, M, N):
    """csrc/qgemm8x.cu k_qgemm8x: eight plain real products of the pre-combined planes (the zero padding contributes
    nothing) and the recombination of qgemm.cu's epilogue.  Returns (Ca, Cb) of the M x N quaternion product."""
    p = [A8[e].T @ B8[e] for e in range(8)]
    s123 = (p[0] + p[1]) + p[2]
    s = 0.5 * (s123 + p[3])
    Ca, Cb = q_from_components([(s - p[0]) + p[4], (s - s123) + p[5], (s - p[1]) + p[6], (s - p[2]) + p[7]])
    return Ca[:M, :N], Cb[:M, :N]


def qgemm_ref(Aa, Ab, Ba, Bb):
    """the same product through the complex 2 x 2 block form Phi(A) (Ba; Bb) (what zgemm.cu's stacked GEMMs compute)"""
    return Aa @ Ba - c(Ab) @ Bb, Ab @ Ba + c(Aa) @ Bb


def backtransform_q8(D, E, tau, Xa, Xb, nb):
    """K6 with the quaternion 8-product GEMMs: per panel Y = V^H X and X -= V (T Y) as quaternion products (T Y stays
    a small stacked complex product, as in solver.cu)."""
    n = D.shape[0]
    Xa, Xb = Xa.copy(), Xb.copy()
    for j0 in reversed(range(0, n - 1, nb)):
        kb = min(nb, n - 1 - j0)
        m = n - 1 - j0
        P = phi_panel(D, E, j0, kb)
        T = tfactor(P, tau[j0:j0 + kb])
        Va, Vb = P[:m, :kb], P[m:, :kb]
        VHa, VHb = c(Va).T, -Vb.T                         # quaternion conjugate transpose in pair form
        Ya, Yb = qgemm8(VHa, VHb, Xa[j0 + 1:], Xb[j0 + 1:])
        TY = T @ np.vstack([Ya, Yb])
        Ua, Ub = qgemm8(Va, Vb, TY[:kb], TY[kb:])
        Xa[j0 + 1:] -= Ua
        Xb[j0 + 1:] -= Ub
    return Xa, Xb
