"""numpy/pure-Python prototype of the device divide-and-conquer solver for the REAL
symmetric tridiagonal eigenproblem (K8).  TEST INFRASTRUCTURE ONLY.

The reference solves its (complex Hermitian) tridiagonal with LAPACK zhbev
(zquatev.cc:84).  The B200 path first makes the tridiagonal real (phase chain) and
then runs Cuppen's divide and conquer on the device: leaf solves, rank-one merges
with deflation, secular equation roots with the Gu-Eisenstat recomputation of z, and
one GEMM per merge.  This file restates that device algorithm step by step (same
array roles and the same order of operations as zquatev_b200/csrc/dc_*.cu) so each
kernel has a CPU answer to be compared with; it is validated against LAPACK
(numpy.linalg.eigh / scipy) in tests/test_oracle.py.
"""
from __future__ import annotations

import numpy as np

EPS = np.finfo(np.float64).eps
LEAF = 32


def tree_levels(n, leaf=LEAF):
    """Balanced splitting: returns list of levels, level[0] = leaves; each level is a list
    of (offset, size).  All leaves have the same depth."""
    segs = [(0, n)]
    levels = [segs]
    while max(s for _, s in segs) > leaf:
        nxt = []
        for off, s in segs:
            h = s // 2
            nxt += [(off, h), (off + h, s - h)]
        segs = nxt
        levels.append(segs)
    return levels[::-1]


def secular_root(j, k, dl, z2, rho):
    """Root j of 1 + rho * sum_i z2[i]/(dl[i]-x) in (dl[j], dl[j+1]) (j = k-1: right of
    dl[k-1]).  Returns (origin index, mu) with root = dl[origin] + mu.  Mirrors the device
    routine `secular_solve` in dc_secular.cuh: origin = nearer pole, bracketed 'middle way'
    rational iteration with bisection safeguard."""
    last = j == k - 1
    if last:
        org = j
        lo, hi = 0.0, rho * float(np.sum(z2))
        Dl = dl - dl[org]
        mu = hi if k == 1 else 0.5 * hi
        if k == 1:
            return org, rho * z2[0]
    else:
        gap = dl[j + 1] - dl[j]
        mid = 0.5 * gap
        Dl = dl - dl[j]
        g = 1.0 + rho * float(np.sum(z2 / (Dl - mid)))
        if g >= 0.0:
            org = j
            lo, hi = 0.0, mid
        else:
            org = j + 1
            Dl = dl - dl[j + 1]
            lo, hi = -mid, 0.0
        mu = 0.5 * (lo + hi)
    pL = Dl[j]
    pR = Dl[j + 1] if not last else 0.0
    for it in range(100):
        t = z2 / (Dl - mu)
        psi = rho * float(np.sum(t[: j + 1]))
        dpsi = rho * float(np.sum(t[: j + 1] / (Dl[: j + 1] - mu)))
        phi = rho * float(np.sum(t[j + 1:]))
        dphi = rho * float(np.sum(t[j + 1:] / (Dl[j + 1:] - mu)))
        g = 1.0 + psi + phi
        err = EPS * (8.0 * (1.0 + abs(psi) + abs(phi)) + abs(mu) * (dpsi + dphi))
        if abs(g) <= err:
            break
        if g < 0.0:
            lo = mu
        else:
            hi = mu
        if not (hi - lo > 2.0 * EPS * max(abs(lo), abs(hi))):
            break
        DL = pL - mu
        new = None
        if last:
            cps = dpsi * DL * DL
            C = 1.0 + psi - dpsi * DL
            if C > 0.0:
                new = mu + DL + cps / C
        else:
            DR = pR - mu
            cps = dpsi * DL * DL
            cph = dphi * DR * DR
            C = 1.0 + (psi - dpsi * DL) + (phi - dphi * DR)
            b = C * (DL + DR) + cps + cph
            cc = DL * DR * g
            # C eta^2 - b eta + cc = 0
            if C == 0.0:
                if b != 0.0:
                    new = mu + cc / b
            else:
                disc = b * b - 4.0 * C * cc
                if disc >= 0.0:
                    sq = np.sqrt(disc)
                    q = 0.5 * (b + (sq if b >= 0 else -sq))
                    cands = []
                    if q != 0.0:
                        cands.append(cc / q)
                    cands.append(q / C)
                    for eta in cands:
                        x = mu + eta
                        if lo < x < hi:
                            new = x
                            break
        if new is None or not (lo < new < hi):
            new = 0.5 * (lo + hi)
        mu = new
    return org, mu


def merge(d, Q, off, n1, n2, rho_in, stats=None, cores=None):
    """One rank-one merge on the diagonal block [off, off+n1+n2) of (d, Q), in place.
    On entry Q[blk] = diag(Q1, Q2) with eigenvalues d (any order inside each child);
    on exit the block holds the merged eigen-decomposition (columns in the order
    [k secular roots ascending | deflated])."""
    nm = n1 + n2
    sl = slice(off, off + nm)
    rho = abs(rho_in)
    sgn = 1.0 if rho_in >= 0 else -1.0
    dd = d[sl].copy()
    Qb = Q[sl, sl]
    z = np.concatenate([Qb[n1 - 1, :n1], sgn * Qb[n1, n1:]]) / np.sqrt(2.0)
    rho = 2.0 * rho
    # rank sort
    order = np.lexsort((np.arange(nm), dd))
    ds = dd[order].copy()
    zs = z[order].copy()
    tol = 8.0 * EPS * max(np.max(np.abs(ds)), np.max(np.abs(zs)))
    nondefl, defl, rots = [], [], []
    if cores is not None:
        return _merge_with_cores(d, Q, sl, Qb, order, ds, zs, rho, nm, stats, cores)
    if rho * np.max(np.abs(zs)) <= tol:
        defl = list(range(nm))
    else:
        pj = -1
        for j in range(nm):
            if rho * abs(zs[j]) <= tol:
                defl.append(j)
                continue
            if pj < 0:
                pj = j
                continue
            s, cc = zs[pj], zs[j]
            tau = np.hypot(cc, s)
            t = ds[j] - ds[pj]
            cc /= tau
            s = -s / tau
            if abs(t * cc * s) <= tol:
                zs[j], zs[pj] = tau, 0.0
                rots.append((order[pj], order[j], cc, s))
                t = ds[pj] * cc * cc + ds[j] * s * s
                ds[j] = ds[pj] * s * s + ds[j] * cc * cc
                ds[pj] = t
                defl.append(pj)
                pj = j
            else:
                nondefl.append(pj)
                pj = j
        if pj >= 0:
            nondefl.append(pj)
    for (c1, c2, cc, s) in rots:
        x, y = Qb[:, c1].copy(), Qb[:, c2].copy()
        Qb[:, c1] = cc * x + s * y
        Qb[:, c2] = cc * y - s * x
    k = len(nondefl)
    if stats is not None:
        stats.append((nm, k))
    Qn = np.zeros_like(Qb)
    dn = np.zeros(nm)
    if k > 0:
        dl = ds[nondefl]
        w = zs[nondefl]
        z2 = w * w
        S = np.zeros((k, k))
        lam = np.zeros(k)
        for j in range(k):
            org, mu = secular_root(j, k, dl, z2, rho)
            S[:, j] = (dl - dl[org]) - mu
            lam[j] = dl[org] + mu
        # Gu-Eisenstat
        zh = np.zeros(k)
        for i in range(k):
            p = S[i, i]
            for jj in range(k):
                if jj != i:
                    p *= S[i, jj] / (dl[i] - dl[jj])
            zh[i] = np.copysign(np.sqrt(abs(p)), w[i])
        S = zh[:, None] / S
        S /= np.linalg.norm(S, axis=0)[None, :]
        Qn[:, :k] = Qb[:, order[nondefl]] @ S
        dn[:k] = lam
    if len(defl):
        Qn[:, k:] = Qb[:, order[defl]]
        dn[k:] = ds[defl]
    Q[sl, sl] = Qn
    d[sl] = dn


def _merge_with_cores(d, Q, sl, Qb, order, ds, zs, rho, nm, stats, cores):
    """Same merge, but deflation scan and secular roots come from the C cores of the device
    code compiled for the host (tests/host_shim) -- used to unit-test those cores on the CPU."""
    k, dl, w, ndcol, dfval, dfcol, rots = cores.deflate(rho, ds, zs, order.astype(np.int32))
    for (c1, c2, cc, s) in rots:
        x, y = Qb[:, c1].copy(), Qb[:, c2].copy()
        Qb[:, c1] = cc * x + s * y
        Qb[:, c2] = cc * y - s * x
    if stats is not None:
        stats.append((nm, k))
    Qn = np.zeros_like(Qb)
    dn = np.zeros(nm)
    if k > 0:
        z2 = w * w
        org, mu, worst = cores.secular(dl, z2, rho)
        if stats is not None:
            stats.append(("iters", worst))
        S = (dl[:, None] - dl[org][None, :]) - mu[None, :]
        lam = dl[org] + mu
        zh = np.zeros(k)
        for i in range(k):
            p = S[i, i]
            for jj in range(k):
                if jj != i:
                    p *= S[i, jj] / (dl[i] - dl[jj])
            zh[i] = np.copysign(np.sqrt(abs(p)), w[i])
        S = zh[:, None] / S
        S /= np.linalg.norm(S, axis=0)[None, :]
        Qn[:, :k] = Qb[:, ndcol] @ S
        dn[:k] = lam
    if nm - k:
        Qn[:, k:] = Qb[:, dfcol]
        dn[k:] = dfval
    Q[sl, sl] = Qn
    d[sl] = dn


def leaf_solve(d, e):
    """Leaf eigen-solve; the device uses implicit-shift QL (dc_leaf.cu)."""
    T = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
    return np.linalg.eigh(T)


def stedc(d_in, e_in, leaf=LEAF, stats=None, cores=None):
    """Eigen-decomposition of the real symmetric tridiagonal (d, e): returns (w ascending, Z)."""
    n = len(d_in)
    d = np.array(d_in, dtype=np.float64)
    e = np.array(e_in, dtype=np.float64)
    if n == 1:
        return d.copy(), np.ones((1, 1))
    scale = max(np.max(np.abs(d)), np.max(np.abs(e)) if n > 1 else 0.0)
    if scale == 0.0:
        return d.copy(), np.eye(n)
    d /= scale
    e /= scale
    levels = tree_levels(n, leaf)
    # tear: every boundary between adjacent leaves
    bounds = sorted({off for lev in levels for off, _ in lev if off > 0})
    for p in bounds:
        d[p - 1] -= abs(e[p - 1])
        d[p] -= abs(e[p - 1])
    Q = np.zeros((n, n))
    for off, s in levels[0]:
        if cores is not None:
            w, Zl = cores.leaf(d[off:off + s], e[off:off + s - 1])
        else:
            w, Zl = leaf_solve(d[off:off + s], e[off:off + s - 1])
        d[off:off + s] = w
        Q[off:off + s, off:off + s] = Zl
    for li in range(1, len(levels)):
        child = levels[li - 1]
        for mi, (off, s) in enumerate(levels[li]):
            (o1, n1), (o2, n2) = child[2 * mi], child[2 * mi + 1]
            assert o1 == off and o2 == off + n1 and n1 + n2 == s
            merge(d, Q, off, n1, n2, e[o2 - 1], stats, cores)
    order = np.lexsort((np.arange(n), d))
    return d[order] * scale, Q[:, order]
