// K8 numerical cores of the on-device divide & conquer for the real symmetric tridiagonal
// eigenproblem (replaces the reference's LAPACK zhbev call, zquatev.cc:84).
//
// Every routine is written once as a __host__ __device__ template over a "lane policy":
//   Lanes::id()      lane index, Lanes::count() lanes cooperating, Lanes::sum(x) all-reduce,
//   Lanes::bcast(x, src)
// On the device the policy is a warp (32 lanes, shuffles); tests/dc_host_shim.cc instantiates
// the same source with a 1-lane policy so the numerics are unit-tested on the CPU against
// LAPACK without a GPU (test infrastructure only -- the product path is the CUDA build).
#pragma once
#include <math.h>
#include <float.h>
#include "common.cuh"

namespace zq {

constexpr double DC_EPS = 2.220446049250313e-16;   // 2^-52
constexpr int DC_LEAF = 32;

struct OneLane {
  ZQ_HD static int id() { return 0; }
  ZQ_HD static int count() { return 1; }
  ZQ_HD static double sum(double x) { return x; }
  ZQ_HD static double maxv(double x) { return x; }
};

#ifdef __CUDACC__
struct WarpLanes {
  ZQ_D static int id() { return threadIdx.x & 31; }
  ZQ_D static int count() { return 32; }
  ZQ_D static double sum(double x) { return warp_sum(x); }
  ZQ_D static double maxv(double x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x = fmax(x, __shfl_xor_sync(0xffffffffu, x, o));
    return x;
  }
};
#endif

// ---------------------------------------------------------------------------------------------
// Secular equation.  Poles dl[0] < ... < dl[k-1], weights z2[i] = z_i^2 > 0, rho > 0:
//     g(x) = 1 + rho * sum_i z2[i] / (dl[i] - x).
// Root j lies in (dl[j], dl[j+1]) (j = k-1: (dl[k-1], dl[k-1] + rho*sum z2)).
// The root is returned as origin index `org` (the nearer pole) plus offset `mu`, so that the
// differences dl[i] - root = (dl[i] - dl[org]) - mu keep full RELATIVE accuracy -- the property
// the Gu-Eisenstat eigenvector formula needs.  Iteration: the "middle way" osculating rational
// interpolation on the two neighbouring poles, kept inside a sign-change bracket with a
// bisection fallback.  Returns the number of iterations (>= maxit: not converged).
// ---------------------------------------------------------------------------------------------
template <class Lanes>
ZQ_HD int secular_root(int j, int k, const double* dl, const double* z2, double rho, int* org_out, double* mu_out) {
  const int lane = Lanes::id(), nl = Lanes::count();
  const int maxit = 100;
  const bool last = (j == k - 1);
  if (k == 1) {
    *org_out = 0;
    *mu_out = rho * z2[0];
    return 0;
  }
  int org;
  double lo, hi;
  if (last) {
    double s = 0.0;
    for (int i = lane; i < k; i += nl) s += z2[i];
    s = Lanes::sum(s);
    org = j;
    lo = 0.0;
    hi = rho * s;
  } else {
    const double mid = 0.5 * (dl[j + 1] - dl[j]);
    const double dj = dl[j];
    double s = 0.0;
    for (int i = lane; i < k; i += nl) s += z2[i] / ((dl[i] - dj) - mid);
    s = Lanes::sum(s);
    const double g = 1.0 + rho * s;
    if (g >= 0.0) { org = j; lo = 0.0; hi = mid; }
    else          { org = j + 1; lo = -mid; hi = 0.0; }
  }
  const double dorg = dl[org];
  const double pL = dl[j] - dorg;
  const double pR = last ? 0.0 : dl[j + 1] - dorg;
  double mu = 0.5 * (lo + hi);
  int it = 0;
  for (; it < maxit; ++it) {
    double psi = 0.0, dpsi = 0.0, phi = 0.0, dphi = 0.0;
    for (int i = lane; i < k; i += nl) {
      const double del = (dl[i] - dorg) - mu;
      const double t = z2[i] / del;
      const double t2 = t / del;
      if (i <= j) { psi += t; dpsi += t2; } else { phi += t; dphi += t2; }
    }
    psi = rho * Lanes::sum(psi);
    dpsi = rho * Lanes::sum(dpsi);
    phi = rho * Lanes::sum(phi);
    dphi = rho * Lanes::sum(dphi);
    const double g = 1.0 + psi + phi;
    const double err = DC_EPS * (8.0 * (1.0 + fabs(psi) + fabs(phi)) + fabs(mu) * (dpsi + dphi));
    if (!(fabs(g) > err)) break;          // also leaves on NaN
    if (g < 0.0) lo = mu; else hi = mu;
    if (!(hi - lo > 2.0 * DC_EPS * fmax(fabs(lo), fabs(hi)))) break;
    const double DL = pL - mu;
    double nw = 0.0;
    bool have = false;
    if (last) {
      const double cps = dpsi * DL * DL;
      const double C = 1.0 + psi - dpsi * DL;
      if (C > 0.0) { nw = mu + DL + cps / C; have = true; }
    } else {
      const double DR = pR - mu;
      const double cps = dpsi * DL * DL, cph = dphi * DR * DR;
      const double C = 1.0 + (psi - dpsi * DL) + (phi - dphi * DR);
      const double b = C * (DL + DR) + cps + cph;
      const double cc = DL * DR * g;
      // C eta^2 - b eta + cc = 0
      if (C == 0.0) {
        if (b != 0.0) { nw = mu + cc / b; have = true; }
      } else {
        const double disc = b * b - 4.0 * C * cc;
        if (disc >= 0.0) {
          const double sq = sqrt(disc);
          const double q = 0.5 * (b + (b >= 0.0 ? sq : -sq));
          if (q != 0.0) {
            const double x1 = mu + cc / q;
            if (x1 > lo && x1 < hi) { nw = x1; have = true; }
          }
          if (!have) {
            const double x2 = mu + q / C;
            if (x2 > lo && x2 < hi) { nw = x2; have = true; }
          }
        }
      }
    }
    if (!have || !(nw > lo && nw < hi)) nw = 0.5 * (lo + hi);
    mu = nw;
  }
  *org_out = org;
  *mu_out = mu;
  return it;
}

// ---------------------------------------------------------------------------------------------
// Leaf solver: implicit-shift QL on a tridiagonal of order m <= DC_LEAF.
// d[m], e[m] (e[m-1] scratch) are overwritten; Z (m x m, leading dim ldz, column-major) must
// hold the identity on entry and returns the eigenvectors: lane r owns rows r, r+nl, ... .
// Returns 0 on success, 1 if some eigenvalue failed to converge in 60 sweeps.
// ---------------------------------------------------------------------------------------------
template <class Lanes>
ZQ_HD int leaf_ql(int m, double* d, double* e, double* Z, int ldz) {
  const int lane = Lanes::id(), nl = Lanes::count();
  const double safmin = DBL_MIN;
  int fail = 0;
  if (m > 0) e[m - 1] = 0.0;
  for (int l = 0; l < m; ++l) {
    int iter = 0;
    for (;;) {
      int mm = l;
      for (; mm < m - 1; ++mm) {
        const double t = fabs(e[mm]);
        if (t * t <= (DC_EPS * fabs(d[mm])) * (DC_EPS * fabs(d[mm + 1])) + safmin) { break; }
      }
      if (mm == l) break;
      if (iter++ == 60) { fail = 1; break; }
      // Wilkinson-type shift from the leading 2x2 of the unreduced block
      double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
      double r = hypot(g, 1.0);
      g = d[mm] - d[l] + e[l] / (g + (g >= 0.0 ? fabs(r) : -fabs(r)));
      double s = 1.0, c = 1.0, p = 0.0;
      int i = mm - 1;
      bool under = false;
      for (; i >= l; --i) {
        double f = s * e[i];
        const double b = c * e[i];
        r = hypot(f, g);
        e[i + 1] = r;
        if (r == 0.0) {            // recover from underflow
          d[i + 1] -= p;
          e[mm] = 0.0;
          under = true;
          break;
        }
        s = f / r;
        c = g / r;
        g = d[i + 1] - p;
        r = (d[i] - g) * s + 2.0 * c * b;
        p = s * r;
        d[i + 1] = g + p;
        g = c * r - b;
        for (int row = lane; row < m; row += nl) {
          double* z0 = Z + row + (size_t)i * ldz;
          double* z1 = Z + row + (size_t)(i + 1) * ldz;
          f = *z1;
          *z1 = s * (*z0) + c * f;
          *z0 = c * (*z0) - s * f;
        }
      }
      if (under) continue;
      d[l] -= p;
      e[l] = g;
      e[mm] = 0.0;
    }
  }
  return fail;
}

// ---------------------------------------------------------------------------------------------
// Deflation scan of one rank-one merge (dlaed2's role).  Inputs sorted ascending by value:
// ds[nm], zs[nm], col[nm] (original column of each sorted entry).  rho > 0.
// Outputs: k and, for the k surviving entries in ascending order, dlam/wz/ndcol; for the nm-k
// deflated entries dfval/dfcol; the list of plane rotations (rc1, rc2, rcs: c, s) to be
// applied IN ORDER to the eigenvector columns.  ndtype[] classifies each surviving column like
// dlaed2 does: 1 = non-zero only in the rows of child 1 (col < n1), 3 = only in the rows of child 2,
// 2 = dense (a rotation mixed columns of both children) -- the merge GEMM skips the zero halves.
// Executed redundantly by every lane (state is scalar); only lane 0 stores.
// ---------------------------------------------------------------------------------------------
struct DeflateOut { int k, ndefl, nrot; };

template <class Lanes>
ZQ_HD DeflateOut deflate_scan(int nm, int n1, double rho, const double* ds, const double* zs, const int* col, double* dlam,
                              double* wz, int* ndcol, int* ndtype, double* dfval, int* dfcol, int* rc1, int* rc2,
                              double* rcc, double* rss) {
  const int lane = Lanes::id(), nl = Lanes::count();
  double dmax = 0.0, zmax = 0.0;
  for (int i = lane; i < nm; i += nl) { dmax = fmax(dmax, fabs(ds[i])); zmax = fmax(zmax, fabs(zs[i])); }
  dmax = Lanes::maxv(dmax);
  zmax = Lanes::maxv(zmax);
  const double tol = 8.0 * DC_EPS * fmax(dmax, zmax);
  DeflateOut o; o.k = 0; o.ndefl = 0; o.nrot = 0;
  const bool st = (lane == 0);
  if (!(rho * zmax > tol)) {       // everything deflates (also taken on NaN)
    for (int i = lane; i < nm; i += nl) { dfval[i] = ds[i]; dfcol[i] = col[i]; }
    o.ndefl = nm;
    return o;
  }
  bool havep = false;
  double pd = 0.0, pz = 0.0;
  int pc = 0, pt = 1;
  for (int j = 0; j < nm; ++j) {
    const double dj = ds[j], zj = zs[j];
    const int cj = col[j];
    if (rho * fabs(zj) <= tol) {
      if (st) { dfval[o.ndefl] = dj; dfcol[o.ndefl] = cj; }
      ++o.ndefl;
      continue;
    }
    const int tj = (cj < n1) ? 1 : 3;
    if (!havep) { havep = true; pd = dj; pz = zj; pc = cj; pt = tj; continue; }
    double s = pz, c = zj;
    const double tau = hypot(c, s);
    const double t = dj - pd;
    c /= tau;
    s = -s / tau;
    if (fabs(t * c * s) <= tol) {
      // rotate (previous, current): z_prev -> 0 (deflates), z_cur -> tau
      if (st) { rc1[o.nrot] = pc; rc2[o.nrot] = cj; rcc[o.nrot] = c; rss[o.nrot] = s; }
      ++o.nrot;
      const double dprev = pd * c * c + dj * s * s;
      const double dcur = pd * s * s + dj * c * c;
      if (st) { dfval[o.ndefl] = dprev; dfcol[o.ndefl] = pc; }
      ++o.ndefl;
      pd = dcur; pz = tau; pc = cj; pt = (pt == tj) ? pt : 2;
    } else {
      if (st) { dlam[o.k] = pd; wz[o.k] = pz; ndcol[o.k] = pc; ndtype[o.k] = pt; }
      ++o.k;
      pd = dj; pz = zj; pc = cj; pt = tj;
    }
  }
  if (havep) {
    if (st) { dlam[o.k] = pd; wz[o.k] = pz; ndcol[o.k] = pc; ndtype[o.k] = pt; }
    ++o.k;
  }
  return o;
}

}  // namespace zq
