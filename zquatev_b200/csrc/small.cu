// K5: whole tridiagonalisation of ONE small matrix (n <= SMALL_N_MAX) in ONE launch of ONE CTA.
//
// The multi-kernel reduction (panel.cu + matvec.cu) pays ~4 dependent launches per column; at
// n = 256 (BASELINE config 5: batches of 2n = 512 Dirac-Fock-style problems) that is ~1000 launches
// of a few microseconds for a matrix whose lower triangles are 1 MB, and the GPU front end, not the
// arithmetic, sets the pace.  Here one CTA keeps the reflector vectors in shared memory and streams
// the L2-resident lower triangles of D and E ONCE per column: the pass that applies the rank-2
// update of column k-1 (unblocked form of reference unblocked.cc:44-131 / LAPACK zhetd2, with one
// quaternion reflector per column as in panel.cu) also multiplies the updated entries with the new
// reflector v_k.  A batch runs one such CTA per lane (stream), so up to 148 problems progress at
// once, one per SM.
//
// Outputs are those of the multi-kernel path (same formulas, so the tridiagonal agrees to rounding):
// d, e, tau, alpha, the reflector tails in the strictly-sub-sub-diagonal parts of D and E, and the
// panel Gram columns G (V[:, 0:i]^H v_i) that the compact-WY T factors of the back-transformation
// are built from (backtransform.cu k_build_T).
#include "kernels.h"

namespace zq {
namespace {

constexpr int S_NT = 256, S_NW = S_NT / 32;

// S_RQ = number of 32-row blocks a lane owns (n <= 32 S_RQ): 4 for n <= 128, 8 up to SMALL_N_MAX
template <int S_RQ>
__global__ void __launch_bounds__(S_NT, 1) k_tridiag_cta(PanelWs w) {
  extern __shared__ quat sm[];
  const int n = w.n, nb = w.nb;
  quat* v = sm;                  // [n] current reflector
  quat* vp = v + n;              // [n] previous reflector
  quat* wv = vp + n;             // [n] w of the previous column (zlatrd's w: p - 1/2 tau (v^H p) v)
  quat* x = wv + n;              // [n] updated column, later p
  quat* yt = x + n;              // [n] transposed sums of the mat-vec
  quat* red = yt + n;            // [S_NW][n] direct partial sums, one row of n per warp
  __shared__ double s_red[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  cplx* const D = w.A;
  cplx* const E = w.A + n;
  const size_t lda = w.lda;

  for (int k = 0; k + 1 < n; ++k) {
    const int s = k + 1;                               // support of v_k: rows [s, n)
    const int j0 = (k / nb) * nb, i = k - j0;          // panel bookkeeping for G only
    // ---- (a) column k of the matrix updated by reflector k-1: d_k and x = M[k+1:, k] ----
    quat cw = qzero(), cv = qzero();
    if (k > 0) { cw = qconj(wv[k]); cv = qconj(vp[k]); }
    double nr[1] = {0.0};
    for (int r = k + tid; r < n; r += S_NT) {
      quat col = qmake(D[(size_t)r + (size_t)k * lda], E[(size_t)r + (size_t)k * lda]);
      if (k > 0) { qfms(col, vp[r], cw); qfms(col, wv[r], cv); }
      if (r == k) {
        w.d[k] = col.a.x;
      } else {
        x[r] = col;
        if (r >= k + 2) nr[0] += qnorm2(col);
      }
    }
    block_sum<1>(nr, s_red);                           // also orders the x[] stores before the reads below
    // ---- (b) reflector (zlarfg analogue, same formulas as k_reflector in panel.cu) ----
    const double rest2 = nr[0];
    const quat x1 = x[s];
    const double x1n2 = qnorm2(x1);
    const double nx2 = rest2 + x1n2;
    quat alpha = qzero(), inv = qzero();
    double tau = 0.0, nx = 0.0;
    if (nx2 > 0.0) {
      nx = sqrt(nx2);
      const double x1n = sqrt(x1n2);
      const quat ph = (x1n > 0.0) ? qscale(x1, 1.0 / x1n) : qmake(cmake(1, 0), cmake(0, 0));
      alpha = qscale(ph, -nx);
      const double u1n = x1n + nx;
      const double u1n2 = u1n * u1n;
      tau = 2.0 * u1n2 / (rest2 + u1n2);
      inv = qscale(qconj(ph), 1.0 / u1n);
    }
    for (int r = s + tid; r < n; r += S_NT) {
      quat vr;
      if (r == s) {
        vr = qmake(cmake(1, 0), cmake(0, 0));
      } else {
        vr = qmul(x[r], inv);
        D[(size_t)r + (size_t)k * lda] = vr.a;         // reflector tail lives where zeros were created
        E[(size_t)r + (size_t)k * lda] = vr.b;
      }
      v[r] = vr;
    }
    if (tid == 0) {
      w.alpha[k] = alpha;
      w.e[k] = nx;
      w.tau[k] = tau;
    }
    __syncthreads();
    // ---- (b2) Gram column of the panel: G[k][t] = V_t^H v_k, t < i (tails of earlier reflectors sit in A) ----
    for (int t = warp; t < i; t += S_NW) {
      const cplx* da = D + (size_t)(j0 + t) * lda;
      const cplx* ea = E + (size_t)(j0 + t) * lda;
      quat acc = qzero();
      cplx dg[S_RQ], eg[S_RQ];
#pragma unroll
      for (int q = 0; q < S_RQ; ++q) {                 // all loads first (predicated), then the arithmetic
        const int r = 32 * q + lane;
        const bool ok = r >= s && r < n;
        dg[q] = ok ? da[r] : cmake(0, 0);
        eg[q] = ok ? ea[r] : cmake(0, 0);
      }
#pragma unroll
      for (int q = 0; q < S_RQ; ++q) {
        const int r = 32 * q + lane;
        if (r >= s && r < n) qfma_cj(acc, qmake(dg[q], eg[q]), v[r]);
      }
      acc = warp_sum(acc);
      if (lane == 0) w.G[(size_t)k * nb + t] = acc;
    }
    // ---- (c) one pass over the lower triangles of rows/cols [s, n): apply the rank-2 update of column k-1,
    //          M -= vp wv^H + wv vp^H, store, and accumulate y = M v with the updated entries ----
    quat acc[S_RQ];
#pragma unroll
    for (int q = 0; q < S_RQ; ++q) acc[q] = qzero();
    for (int c = s + warp; c < n; c += S_NW) {
      const quat vc = v[c];
      quat cwc = qzero(), cvc = qzero();
      if (k > 0) { cwc = qconj(wv[c]); cvc = qconj(vp[c]); }
      quat tacc = qzero();
      // every load of the column is issued before the first dependent instruction (predicated, no branches): a
      // branchy per-block form serialises up to S_RQ L2 round trips per column
      cplx dv[S_RQ], ev[S_RQ];
#pragma unroll
      for (int q = 0; q < S_RQ; ++q) {
        const int r = 32 * q + lane;
        const bool ok = r >= c && r < n;
        dv[q] = ok ? D[(size_t)r + (size_t)c * lda] : cmake(0, 0);
        ev[q] = (ok && r > c) ? E[(size_t)r + (size_t)c * lda] : cmake(0, 0);
      }
#pragma unroll
      for (int q = 0; q < S_RQ; ++q) {
        const int r = 32 * q + lane;
        if (r >= c && r < n) {
          quat m = qmake(dv[q], ev[q]);
          if (k > 0) {
            qfms(m, vp[r], cwc);
            qfms(m, wv[r], cvc);
            if (r == c) { m.a.y = 0.0; m.b = cmake(0, 0); }
            D[(size_t)r + (size_t)c * lda] = m.a;
            if (r > c) E[(size_t)r + (size_t)c * lda] = m.b;
          }
          if (r == c) {                                // diagonal: D real, E zero; direct term only
            acc[q].a.x = fma(m.a.x, vc.a.x, acc[q].a.x); acc[q].a.y = fma(m.a.x, vc.a.y, acc[q].a.y);
            acc[q].b.x = fma(m.a.x, vc.b.x, acc[q].b.x); acc[q].b.y = fma(m.a.x, vc.b.y, acc[q].b.y);
          } else {
            const cplx d = m.a, e = m.b;
            const quat vr = v[r];
            // direct: ya[r] += d va[c] - conj(e) vb[c] ; yb[r] += e va[c] + conj(d) vb[c]
            cfma(acc[q].a, d, vc.a);  cfms_ca(acc[q].a, e, vc.b);
            cfma(acc[q].b, e, vc.a);  cfma_ca(acc[q].b, d, vc.b);
            // transposed: ya[c] += conj(d) va[r] + conj(e) vb[r] ; yb[c] += -e va[r] + d vb[r]
            cfma_ca(tacc.a, d, vr.a);  cfma_ca(tacc.a, e, vr.b);
            cfms(tacc.b, e, vr.a);     cfma(tacc.b, d, vr.b);
          }
        }
      }
      tacc = warp_sum(tacc);
      if (lane == 0) yt[c] = tacc;
    }
#pragma unroll
    for (int q = 0; q < S_RQ; ++q) {
      const int r = 32 * q + lane;
      if (r < n) red[(size_t)warp * n + r] = acc[q];
    }
    __syncthreads();
    // ---- (d) p = tau (M v), g = Re(v^H p), w = p - 1/2 tau g v ----
    double g[1] = {0.0};
    for (int r = s + tid; r < n; r += S_NT) {
      quat y = yt[r];
#pragma unroll
      for (int wi = 0; wi < S_NW; ++wi) y = qadd(y, red[(size_t)wi * n + r]);
      y = qscale(y, tau);
      x[r] = y;
      const quat f = v[r];
      g[0] += f.a.x * y.a.x + f.a.y * y.a.y + f.b.x * y.b.x + f.b.y * y.b.y;
    }
    block_sum<1>(g, s_red);
    const double coef = 0.5 * tau * g[0];
    for (int r = s + tid; r < n; r += S_NT) {
      const quat pr = x[r], vr = v[r];
      wv[r] = qmake(csub(pr.a, cscale(vr.a, coef)), csub(pr.b, cscale(vr.b, coef)));
      vp[r] = vr;
    }
    __syncthreads();
  }
  // last diagonal entry: D[n-1, n-1] updated by the last reflector pair
  if (tid == 0) {
    quat col = qmake(D[(size_t)(n - 1) + (size_t)(n - 1) * lda], cmake(0, 0));
    if (n > 1) {
      qfms(col, vp[n - 1], qconj(wv[n - 1]));
      qfms(col, wv[n - 1], qconj(vp[n - 1]));
    }
    w.d[n - 1] = col.a.x;
  }
}

size_t small_smem(int n) { return (size_t)(5 + S_NW) * (size_t)n * sizeof(quat); }

}  // namespace

int small_n_max() {
  const char* e = getenv("ZQ_SMALL_N");
  int v = e ? atoi(e) : SMALL_N_MAX;
  if (v > SMALL_N_MAX) v = SMALL_N_MAX;
  return v < 0 ? 0 : v;
}

cudaError_t small_prepare() {
  cudaError_t e = cudaFuncSetAttribute(k_tridiag_cta<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)small_smem(SMALL_N_MAX));
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(k_tridiag_cta<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)small_smem(128));
  return e;
}

void launch_tridiag_small(const PanelWs& w, cudaStream_t st) {
  if (w.n <= 128) k_tridiag_cta<4><<<1, S_NT, small_smem(w.n), st>>>(w);
  else            k_tridiag_cta<8><<<1, S_NT, small_smem(w.n), st>>>(w);
}

}  // namespace zq
