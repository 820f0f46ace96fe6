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

// Tiles are launched as a 2-D grid (x = row block, y = column block).  x runs fastest, so the CTAs
// resident at any moment stream long contiguous runs of the same 64 columns instead of 2 KB
// pieces of every column (DRAM-page friendly); grid cells above the diagonal exit at once.
inline int count_tiles(int s, int n) {
  const int I0 = s / TR, I1 = (n - 1) / TR, J0 = s / TC, Jlast = (n - 1) / TC;
  int tot = 0;
  for (int I = I0; I <= I1; ++I) tot += (2 * I + 1 < Jlast ? 2 * I + 1 : Jlast) - J0 + 1;
  return tot;
}

// streaming load with an L2 eviction-priority hint (createpolicy evict_first): the 4.3 GB a large mat-vec streams then leave
// the L2 before the ~70 MB panel (V, W) that col_update / the dot CTAs / reduce_correct re-read at every column
// (ZQ_K1_EVICT=1; the plain 16-byte load has no direct .L2::evict_first form -- ptxas accepts that only on 256-bit loads)
ZQ_D unsigned long long l2_evict_first_policy() {
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
ZQ_D cplx ld_stream_hint(const cplx* p, unsigned long long pol) {
  cplx v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(pol));
  return v;
}
template <bool EVF>
ZQ_D cplx ld_mat(const cplx* p, unsigned long long pol) { return EVF ? ld_stream_hint(p, pol) : ld_stream(p); }

template <bool FULL, bool EVF>
ZQ_D void tile_body(const cplx* __restrict__ A, size_t lda, int n, int r0, int c0, const quat (&vrow)[RI],
                    const quat* vcol, quat (&acc)[RI], quat* pt_row, int lane, int warp, unsigned long long pol) {
  const cplx* Dp = A + (size_t)r0 + lane;
  const cplx* Ep = Dp + n;
#pragma unroll 2
  for (int jj = 0; jj < CW; ++jj) {
    const int c = c0 + warp * CW + jj;
    cplx dv[RI], ev[RI];
    if (FULL) {
#pragma unroll
      for (int i = 0; i < RI; ++i) {
        dv[i] = ld_mat<EVF>(Dp + (size_t)c * lda + 32 * i, pol);
        ev[i] = ld_mat<EVF>(Ep + (size_t)c * lda + 32 * i, pol);
      }
    } else {
#pragma unroll
      for (int i = 0; i < RI; ++i) {
        const int r = r0 + lane + 32 * i;
        const bool ok = (r < n) && (c < n) && (r >= c);
        dv[i] = ok ? ld_mat<EVF>(Dp + (size_t)c * lda + 32 * i, pol) : cmake(0, 0);
        ev[i] = (ok && r > c) ? ld_mat<EVF>(Ep + (size_t)c * lda + 32 * i, pol) : cmake(0, 0);
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

// One CTA = one row tile I (128 rows) x one RUN of up to `tpb` owned column blocks (64 columns each, consecutive in
// the owner's numbering: J = rank + (R*tpb + t)*G).  The direct sums of the run accumulate in registers and are
// reduced across the 8 warps once per run (pd[R][r]: 1/tpb of the partial-sum traffic reduce_correct has to read
// back, and 1/tpb of the per-CTA prologue / epilogue); the transposed sums go to pt[I][c] per tile as before.
constexpr int TPB_MAX = K1_TPB_MAX;

template <bool EVF>
__global__ void __launch_bounds__(256, 2)
k_matvec(const cplx* __restrict__ A, size_t lda, int n, int s, const quat* x, int xrec, int head, quat* __restrict__ pd,
         quat* __restrict__ pt, int nI, int R0, int nR, int G, int rank, int tpb, int rev,
         // fused panel dots
         const cplx* __restrict__ pan, int nb, int ncols, int nch, int crows, quat* __restrict__ dotW, quat* __restrict__ dotV) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // dynamic shared memory: red[NW][TR] (32 KB) | vrow_s[TR] (4 KB) | vcol[tpb * TC] (2 KB per column block)
  extern __shared__ __align__(16) unsigned char k1_smem[];
  quat (*red)[TR] = reinterpret_cast<quat (*)[TR]>(k1_smem);
  quat* vrow_s = reinterpret_cast<quat*>(k1_smem) + NW * TR;
  quat* vcol = vrow_s + TR;
  pdl_enter();
  const unsigned long long pol = EVF ? l2_evict_first_policy() : 0ull;
  const quat inv = ld_cg_quat(x + xrec + 2);
  if ((int)blockIdx.x >= nI) {
    // ---- panel inner products: the grid cells right of the tile columns, flattened, one per (row chunk, group of
    // 8 panel columns); warp = one panel column, lanes stride the chunk's rows.  v of the chunk is staged in shared
    // memory once per CTA (x is read at L2; eight warps re-reading it there was measurable) ----
    const int id = ((int)blockIdx.x - nI) + ((int)gridDim.x - nI) * (int)blockIdx.y;
    const int ch = id % nch, t = (id / nch) * NW + warp;
    if ((id / nch) * NW >= ncols) return;                    // whole CTA idle (uniform): before any barrier
    const int ra = s + ch * crows, rb = min(n, ra + crows);
    const bool mine = t < ncols;
    const int tt = mine ? t : 0;
    const cplx* va = pan + ((size_t)(0 * nb + tt)) * n;
    const cplx* vb = pan + ((size_t)(1 * nb + tt)) * n;
    const cplx* wa = pan + ((size_t)(2 * nb + tt)) * n;
    const cplx* wb = pan + ((size_t)(3 * nb + tt)) * n;
    quat* fs = &red[0][0];                                   // NW * TR = 1024 staged rows
    quat aW = qzero(), aV = qzero();
    for (int base = ra; base < rb; base += NW * TR) {
      const int cnt = min(NW * TR, rb - base);
      __syncthreads();
      for (int e = threadIdx.x; e < cnt; e += 256) fs[e] = refl_v(x, inv, base + e, s, head, n);
      __syncthreads();
      if (mine) {
#pragma unroll 4
        for (int e = lane; e < cnt; e += 32) {
          const quat f = fs[e];
          const int r = base + e;
          qfma_cj(aW, qmake(wa[r], wb[r]), f);
          qfma_cj(aV, qmake(va[r], vb[r]), f);
        }
      }
    }
    aW = warp_sum(aW);
    aV = warp_sum(aV);
    if (lane == 0 && mine) {
      dotW[(size_t)ch * nb + t] = aW;
      dotV[(size_t)ch * nb + t] = aV;
    }
    return;
  }
  // rev: sweep the tile grid backwards.  Consecutive columns alternate the direction, so the ~100 MB of D and E that
  // the previous sweep touched last are still in the 126 MB L2 when this one starts there (and the first sweep after
  // a trailing update starts where the GEMM wrote last).  Tile -> partial-buffer mapping is unchanged: same sums.
  const int I = s / TR + (rev ? nI - 1 - (int)blockIdx.x : (int)blockIdx.x);
  const int R = R0 + (rev ? nR - 1 - (int)blockIdx.y : (int)blockIdx.y);
  const int J0 = s / TC, Jlast = (n - 1) / TC;
  const int Jmax = min(2 * I + 1, Jlast);                    // last column block that touches row tile I
  const int Jrun0 = rank + R * tpb * G;                      // first block of the run (owner's numbering)
  if (Jrun0 > Jmax || Jrun0 + (tpb - 1) * G < J0) return;    // run entirely above the diagonal / left of the trailing matrix
  const int r0 = I * TR;
  // stage v: the 128 rows of the tile and the (up to) tpb x 64 columns of the run
  if (threadIdx.x < TR) vrow_s[threadIdx.x] = refl_v(x, inv, r0 + threadIdx.x, s, head, n);
  for (int e = threadIdx.x; e < tpb * TC; e += 256) {
    const int J = Jrun0 + (e / TC) * G;
    vcol[e] = (J >= J0 && J <= Jmax) ? refl_v(x, inv, J * TC + (e % TC), s, head, n) : qzero();
  }
  __syncthreads();
  quat vrow[RI], acc[RI];
#pragma unroll
  for (int i = 0; i < RI; ++i) {
    vrow[i] = vrow_s[lane + 32 * i];
    acc[i] = qzero();
  }
  quat* pt_row = pt + (size_t)I * n;
  for (int t = 0; t < tpb; ++t) {
    const int J = Jrun0 + t * G;
    if (J < J0) continue;
    if (J > Jmax) break;
    const int c0 = J * TC;
    const bool full = (c0 + TC - 1 < r0) && (r0 + TR <= n);
    if (full) tile_body<true, EVF>(A, lda, n, r0, c0, vrow, vcol + t * TC, acc, pt_row, lane, warp, pol);
    else      tile_body<false, EVF>(A, lda, n, r0, c0, vrow, vcol + t * TC, acc, pt_row, lane, warp, pol);
  }
#pragma unroll
  for (int i = 0; i < RI; ++i) red[warp][lane + 32 * i] = acc[i];
  __syncthreads();
  if (threadIdx.x < TR) {
    const int r = r0 + threadIdx.x;
    quat sum = red[0][threadIdx.x];
#pragma unroll
    for (int wv = 1; wv < NW; ++wv) sum = qadd(sum, red[wv][threadIdx.x]);
    if (r < n) pd[(size_t)R * n + r] = sum;
  }
}

// y[r] = sum of partials, rows [s, n) -- only used by the stand-alone test/bench entry (one GPU; tpb as launched)
__global__ void k_matvec_gather(int n, int s, int tpb, const quat* pd, const quat* pt, quat* y) {
  const int r = s + blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const int J0 = s / TC, Jlast = (n - 1) / TC, I0 = s / TR, I1 = (n - 1) / TR;
  const int Jhi = min(2 * (r / TR) + 1, Jlast), Ilo = max(I0, (r / TC) / 2);
  quat acc = qzero();
  for (int R = J0 / tpb; R <= Jhi / tpb; ++R) acc = qadd(acc, pd[(size_t)R * n + r]);
  for (int I = Ilo; I <= I1; ++I) acc = qadd(acc, pt[(size_t)I * n + r]);
  y[r] = acc;
}

}  // namespace

// Column blocks per CTA from the trailing size (measured, profiles/r02_k1_sweep.jsonl: fraction of the copy bandwidth at
// m = 16384 / 8192 / 4096 is 0.83 / 0.81 / 0.67 with 1, 0.89 / 0.84 / 0.75 with 2, 0.93 / 0.85 / 0.75 with 4 blocks;
// below m ~ 3500 the grid gets too small for runs).  ZQ_K1_TPB forces a value; ZQ_K1_TPB8 = trailing size from which 8
// blocks are used.
int k1_tpb(int m, int world) {
  static const int env = [] { const char* e = getenv("ZQ_K1_TPB"); return e ? atoi(e) : 0; }();
  static const int m8 = [] { const char* e = getenv("ZQ_K1_TPB8"); return e ? atoi(e) : 1 << 30; }();
  if (env > 0) return env < K1_TPB_MAX ? env : K1_TPB_MAX;
  // with the column blocks dealt out to `world` ranks a launch holds 1/world of the tiles: the same number of CTAs as a
  // single-GPU launch of trailing size m / sqrt(world)
  if (world > 1) m = (int)((double)m / sqrt((double)world));
  if (m >= m8) return 8;
  return m >= 7168 ? 4 : (m >= 3584 ? 2 : 1);
}

static size_t k1_smem_bytes(int tpb) { return (size_t)(NW * TR + TR + tpb * TC) * sizeof(quat); }
static void k1_prepare() {
  static std::atomic<unsigned long long> done{0};
  if (first_use_on_this_device(done)) {
    cudaFuncSetAttribute(k_matvec<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k1_smem_bytes(K1_TPB_MAX));
    cudaFuncSetAttribute(k_matvec<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k1_smem_bytes(K1_TPB_MAX));
  }
}
// ZQ_K1_EVICT: trailing size from which the matrix loads carry the L2 evict_first hint (0 = never; read at every solve)
static bool k1_evict(int m) {
  const char* e = getenv("ZQ_K1_EVICT");
  const int from = e ? atoi(e) : 0;
  return from > 0 && m >= from;
}

// runs of this rank that hold at least one column block J >= J0: first index R0 and count
static void owned_runs(const PanelWs& w, int s, int tpb, int& R0, int& nR) {
  const int J0 = s / TC, Jlast = (w.n - 1) / TC, G = w.world;
  const int jfirst = J0 + ((w.rank - J0 % G) + G) % G;      // first owned block >= J0
  if (jfirst > Jlast) { R0 = 0; nR = 0; return; }
  const int qfirst = jfirst / G, qlast = (Jlast - w.rank) / G;
  R0 = qfirst / tpb;
  nR = qlast / tpb - R0 + 1;
}

void launch_matvec(const PanelWs& w, int k, int j0, cudaStream_t st) {
  const int s = k + 1, n = w.n;
  const int nI = (n - 1) / TR - s / TR + 1;
  const int tpb = k1_tpb(n - s, w.world);
  int R0, nR;
  owned_runs(w, s, tpb, R0, nR);
  const int ncols = k - j0;
  const int crows = dot_chunk_rows(n - s);
  const int nch = ncols > 0 ? (n - s + crows - 1) / crows : 0;
  const int ndot = nch * ((ncols + NW - 1) / NW);       // (row chunk, group of 8 panel columns) cells
  if (nR == 0 && ndot == 0) return;
  static const int zigzag = [] { const char* e = getenv("ZQ_K1_ZIGZAG"); return e ? atoi(e) : 1; }();
  const int rev = (zigzag && nR > 0 && ((k - j0) & 1) == 0) ? 1 : 0;
  const int gy = nR > 0 ? nR : 1;
  k1_prepare();
  launch_chain_smem(k1_evict(n - s) ? k_matvec<true> : k_matvec<false>, dim3(nI + (ndot + gy - 1) / gy, gy), dim3(256), k1_smem_bytes(tpb), st, w.A, w.lda, n, s,
                    (const quat*)w.x, w.xrec, s, w.pd, w.pt, nI, R0, nR, w.world, w.rank, tpb, rev, w.pan, w.nb, ncols, nch > 0 ? nch : 1, crows, w.dotW,
                    w.dotV);
}

void launch_matvec_only(const PanelWs& w, int s, quat* y, cudaStream_t st, bool gather) {
  const int n = w.n;
  const int nI = (n - 1) / TR - s / TR + 1;
  const int tpb = k1_tpb(n - s, 1);
  int R0, nR;
  owned_runs(w, s, tpb, R0, nR);
  // head = -1: no unit head row, v = x * record[2] for all rows >= s
  k1_prepare();
  (k1_evict(n - s) ? k_matvec<true> : k_matvec<false>)<<<dim3(nI, nR), 256, k1_smem_bytes(tpb), st>>>(w.A, w.lda, n, s, (const quat*)w.x, w.xrec, -1, w.pd, w.pt, nI, R0, nR, 1, 0, tpb, 0, w.pan, w.nb, 0, 1,
                                         DOT_MIN_ROWS, w.dotW, w.dotV);
  if (gather) k_matvec_gather<<<(n - s + 255) / 256, 256, 0, st>>>(n, s, tpb, w.pd, w.pt, y);
}

}  // namespace zq
