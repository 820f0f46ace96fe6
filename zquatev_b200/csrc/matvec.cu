// K1: quaternion-Hermitian matrix-vector product on the LOWER triangles,
//       y = (D + jE) v  restricted to rows/cols [s, n),
// the HBM-bound kernel of the tridiagonalisation.  It replaces the four zgemv passes per
// column of the reference (blocked.cc:214,219,397,402 -- each streams the trailing D0 or D1
// once) by ONE pass over the lower triangles of D and E: 16*m^2 bytes per column instead of
// the reference's 64*m^2.
//
// Tile = 128 rows x 64 columns, one CTA of 8 warps.  Warp w owns 8 columns of the tile, lane l
// owns rows l, l+32, l+64, l+96, so every load instruction of a warp reads 512 contiguous bytes
// (32 consecutive complex<double> of one column).  Each loaded pair (d,e) = (D[r,c], E[r,c])
// feeds both the direct product (row r, accumulated in registers over the warp's columns) and
// the transposed product (column c, accumulated over the lane's 4 rows and then reduced with
// warp shuffles).  Direct partials are summed across the 8 warps through shared memory and
// written to pd[J][r]; transposed sums go to pt[I][c]; reduce_correct (panel.cu) adds them in
// a fixed order -> bit-reproducible, no atomics.
//
// The same launch carries extra CTAs that compute the partial panel inner products
// W^H v and V^H v (skinny GEMV^H over rows [s, n)), which the correction step needs: one CTA per
// (row chunk, group of 8 panel columns), see dot_chunk_rows (kernels.h).
#include "kernels.h"

namespace zq {
namespace {

constexpr int TR = MV_TR, TC = MV_TC, RI = TR / 32, NW = 8, CW = TC / NW;

struct TileIdx { int I, J; };

// Tiles are launched as a 2-D grid (x = row block, y = column block).  x runs fastest, so the CTAs
// resident at any moment stream long contiguous runs of the same 64 columns instead of 2 KB
// pieces of every column (DRAM-page friendly); grid cells above the diagonal exit at once.
inline int count_tiles(int s, int n) {
  const int I0 = s / TR, I1 = (n - 1) / TR, J0 = s / TC, Jlast = (n - 1) / TC;
  int tot = 0;
  for (int I = I0; I <= I1; ++I) tot += (2 * I + 1 < Jlast ? 2 * I + 1 : Jlast) - J0 + 1;
  return tot;
}

template <bool FULL>
ZQ_D void tile_body(const cplx* __restrict__ A, size_t lda, int n, int r0, int c0, const quat (&vrow)[RI],
                    const quat* vcol, quat (&acc)[RI], quat* pt_row, int lane, int warp) {
  const cplx* Dp = A + (size_t)r0 + lane;
  const cplx* Ep = Dp + n;
#pragma unroll 2
  for (int jj = 0; jj < CW; ++jj) {
    const int c = c0 + warp * CW + jj;
    cplx dv[RI], ev[RI];
    if (FULL) {
#pragma unroll
      for (int i = 0; i < RI; ++i) {
        dv[i] = ld_stream(Dp + (size_t)c * lda + 32 * i);
        ev[i] = ld_stream(Ep + (size_t)c * lda + 32 * i);
      }
    } else {
#pragma unroll
      for (int i = 0; i < RI; ++i) {
        const int r = r0 + lane + 32 * i;
        const bool ok = (r < n) && (c < n) && (r >= c);
        dv[i] = ok ? ld_stream(Dp + (size_t)c * lda + 32 * i) : cmake(0, 0);
        ev[i] = (ok && r > c) ? ld_stream(Ep + (size_t)c * lda + 32 * i) : cmake(0, 0);
      }
    }
    const quat vc = vcol[warp * CW + jj];
    quat tacc = qzero();
#pragma unroll
    for (int i = 0; i < RI; ++i) {
      cplx d = dv[i];
      const cplx e = ev[i];
      if (!FULL) {
        const int r = r0 + lane + 32 * i;
        if (r == c) {                       // diagonal: D real, E zero; direct term only
          acc[i].a.x = fma(d.x, vc.a.x, acc[i].a.x); acc[i].a.y = fma(d.x, vc.a.y, acc[i].a.y);
          acc[i].b.x = fma(d.x, vc.b.x, acc[i].b.x); acc[i].b.y = fma(d.x, vc.b.y, acc[i].b.y);
          d = cmake(0, 0);
        }
      }
      // direct: ya[r] += d va[c] - conj(e) vb[c] ; yb[r] += e va[c] + conj(d) vb[c]
      cfma(acc[i].a, d, vc.a);  cfms_ca(acc[i].a, e, vc.b);
      cfma(acc[i].b, e, vc.a);  cfma_ca(acc[i].b, d, vc.b);
      // transposed: ya[c] += conj(d) va[r] + conj(e) vb[r] ; yb[c] += -e va[r] + d vb[r]
      cfma_ca(tacc.a, d, vrow[i].a);  cfma_ca(tacc.a, e, vrow[i].b);
      cfms(tacc.b, e, vrow[i].a);     cfma(tacc.b, d, vrow[i].b);
    }
    tacc = warp_sum(tacc);
    if (lane == 0 && c < n) pt_row[c] = tacc;
  }
}

// v[r] of the current reflector: 1 at the head row, x[r] u1^{-1} below it, 0 above (rows of a tile that lie above
// the trailing matrix).  x and the record are read at L2 (.cg): on several GPUs they were stored by a peer, and on
// one GPU the record is written by the last CTA of the previous kernel.
ZQ_D quat ld_cg_quat(const quat* p) {
  const double2* q = reinterpret_cast<const double2*>(p);
  return qmake(__ldcg(q), __ldcg(q + 1));
}
ZQ_D quat refl_v(const quat* x, quat inv, int r, int s, int head, int n) {
  if (r >= n || r < s) return qzero();
  if (r == head) return qmake(cmake(1, 0), cmake(0, 0));
  return qmul(ld_cg_quat(x + r), inv);
}

__global__ void __launch_bounds__(256, 2)
k_matvec(const cplx* __restrict__ A, size_t lda, int n, int s, const quat* x, int xrec, int head, quat* __restrict__ pd,
         quat* __restrict__ pt, int nI, int jfirst, int jstride, int nJ, int rev,
         // fused panel dots
         const cplx* __restrict__ pan, int nb, int ncols, int nch, int crows, quat* __restrict__ dotW, quat* __restrict__ dotV) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  pdl_enter();
  const quat inv = ld_cg_quat(x + xrec + 2);
  if ((int)blockIdx.x >= nI) {
    // ---- panel inner products: the grid cells right of the tile columns, flattened, one per (row chunk, group of
    // 8 panel columns); warp = one panel column, lanes stride the chunk's rows ----
    const int id = ((int)blockIdx.x - nI) + ((int)gridDim.x - nI) * (int)blockIdx.y;
    const int ch = id % nch, t = (id / nch) * NW + warp;
    if (t >= ncols) return;
    const int ra = s + ch * crows, rb = min(n, ra + crows);
    const cplx* va = pan + ((size_t)(0 * nb + t)) * n;
    const cplx* vb = pan + ((size_t)(1 * nb + t)) * n;
    const cplx* wa = pan + ((size_t)(2 * nb + t)) * n;
    const cplx* wb = pan + ((size_t)(3 * nb + t)) * n;
    quat aW = qzero(), aV = qzero();
#pragma unroll 4
    for (int r = ra + lane; r < rb; r += 32) {
      const quat f = refl_v(x, inv, r, s, head, n);
      qfma_cj(aW, qmake(wa[r], wb[r]), f);
      qfma_cj(aV, qmake(va[r], vb[r]), f);
    }
    aW = warp_sum(aW);
    aV = warp_sum(aV);
    if (lane == 0) {
      dotW[(size_t)ch * nb + t] = aW;
      dotV[(size_t)ch * nb + t] = aV;
    }
    return;
  }
  __shared__ quat vcol[TC];
  __shared__ quat red[NW][TR];
  TileIdx ti;
  // rev: sweep the tile grid backwards.  Consecutive columns alternate the direction, so the ~100 MB of D and E that
  // the previous sweep touched last are still in the 126 MB L2 when this one starts there (and the first sweep after
  // a trailing update starts where the GEMM wrote last).  Tile -> partial-buffer mapping is unchanged: same sums.
  ti.I = s / TR + (rev ? nI - 1 - (int)blockIdx.x : (int)blockIdx.x);
  ti.J = jfirst + (rev ? nJ - 1 - (int)blockIdx.y : (int)blockIdx.y) * jstride;   // owned column blocks only (multi-GPU)
  if (2 * ti.I + 1 < ti.J || ti.J * TC >= n) return;    // tile entirely above the diagonal / no owned block
  const int r0 = ti.I * TR, c0 = ti.J * TC;
  if (threadIdx.x < TC) {
    const int c = c0 + threadIdx.x;
    vcol[threadIdx.x] = refl_v(x, inv, c, s, head, n);
  }
  quat vrow[RI], acc[RI];
#pragma unroll
  for (int i = 0; i < RI; ++i) {
    const int r = r0 + lane + 32 * i;
    vrow[i] = refl_v(x, inv, r, s, head, n);
    acc[i] = qzero();
  }
  __syncthreads();
  const bool full = (c0 + TC - 1 < r0) && (r0 + TR <= n);
  quat* pt_row = pt + (size_t)ti.I * n;
  if (full) tile_body<true>(A, lda, n, r0, c0, vrow, vcol, acc, pt_row, lane, warp);
  else      tile_body<false>(A, lda, n, r0, c0, vrow, vcol, acc, pt_row, lane, warp);
#pragma unroll
  for (int i = 0; i < RI; ++i) red[warp][lane + 32 * i] = acc[i];
  __syncthreads();
  if (threadIdx.x < TR) {
    const int r = r0 + threadIdx.x;
    quat sum = red[0][threadIdx.x];
#pragma unroll
    for (int wv = 1; wv < NW; ++wv) sum = qadd(sum, red[wv][threadIdx.x]);
    if (r < n) pd[(size_t)ti.J * n + r] = sum;
  }
}

// y[r] = sum of partials, rows [s, n) -- only used by the stand-alone test/bench entry
__global__ void k_matvec_gather(int n, int s, const quat* pd, const quat* pt, quat* y) {
  const int r = s + blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const int J0 = s / TC, Jlast = (n - 1) / TC, I0 = s / TR, I1 = (n - 1) / TR;
  const int Jhi = min(2 * (r / TR) + 1, Jlast), Ilo = max(I0, (r / TC) / 2);
  quat acc = qzero();
  for (int J = J0; J <= Jhi; ++J) acc = qadd(acc, pd[(size_t)J * n + r]);
  for (int I = Ilo; I <= I1; ++I) acc = qadd(acc, pt[(size_t)I * n + r]);
  y[r] = acc;
}

}  // namespace

static void owned_blocks(const PanelWs& w, int s, int& jfirst, int& nJ) {
  const int J0 = s / TC, Jlast = (w.n - 1) / TC;
  jfirst = J0 + ((w.rank - J0 % w.world) + w.world) % w.world;
  nJ = jfirst > Jlast ? 0 : (Jlast - jfirst) / w.world + 1;
}

void launch_matvec(const PanelWs& w, int k, int j0, cudaStream_t st) {
  const int s = k + 1, n = w.n;
  const int nI = (n - 1) / TR - s / TR + 1;
  int jfirst, nJ;
  owned_blocks(w, s, jfirst, nJ);
  const int ncols = k - j0;
  const int crows = dot_chunk_rows(n - s);
  const int nch = ncols > 0 ? (n - s + crows - 1) / crows : 0;
  const int ndot = nch * ((ncols + NW - 1) / NW);       // (row chunk, group of 8 panel columns) cells
  if (nJ == 0 && ndot == 0) return;
  static const int zigzag = [] { const char* e = getenv("ZQ_K1_ZIGZAG"); return e ? atoi(e) : 1; }();
  const int rev = (zigzag && nJ > 0 && ((k - j0) & 1) == 0) ? 1 : 0;
  const int gy = nJ > 0 ? nJ : 1;
  launch_chain(k_matvec, dim3(nI + (ndot + gy - 1) / gy, gy), dim3(256), st, w.A, w.lda, n, s, (const quat*)w.x, w.xrec, s, w.pd, w.pt, nI, jfirst, w.world, gy,
               rev, w.pan, w.nb, ncols, nch > 0 ? nch : 1, crows, w.dotW, w.dotV);
}

void launch_matvec_only(const PanelWs& w, int s, quat* y, cudaStream_t st) {
  const int n = w.n;
  const int nI = (n - 1) / TR - s / TR + 1, nJ = (n - 1) / TC - s / TC + 1;
  // head = -1: no unit head row, v = x * record[2] for all rows >= s
  k_matvec<<<dim3(nI, nJ), 256, 0, st>>>(w.A, w.lda, n, s, (const quat*)w.x, w.xrec, -1, w.pd, w.pt, nI, s / TC, 1, nJ, 0, w.pan, w.nb, 0, 1, DOT_MIN_ROWS, w.dotW, w.dotV);
  k_matvec_gather<<<(n - s + 255) / 256, 256, 0, st>>>(n, s, w.pd, w.pt, y);
}

}  // namespace zq
