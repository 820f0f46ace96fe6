// K8: divide & conquer eigensolver for the REAL symmetric tridiagonal (d, e), entirely on
// the device.  Replaces LAPACK zhbev (zhbtrd + zsteqr, a serial O(n^3) rotation loop that is
// 63-76 % of the reference's run time: SURVEY.md 3.2) called at zquatev.cc:84.
//
// Balanced tree over the index range, leaves of order <= 32 solved by implicit QL (one warp
// each); every level is ONE batch of kernels over all merges of that level:
//   prep      z = (last row of Q1 | first row of Q2)/sqrt2, rank-sort of the children's eigenvalues
//   deflate   sequential deflation scan, one warp per merge (dc_core.cuh: deflate_scan)
//   rotate    apply the deflation plane rotations to the eigenvector columns
//   secular   one warp per root (dc_core.cuh: secular_root), writes the matrix dl_i - lambda_j
//   zhat      Gu-Eisenstat recomputation of z from the computed roots
//   vectors   eigenvectors of the rank-one update, normalised
//   gemm      Qnew[:, 0:k] = Qold[:, surviving columns] * S   (FP64 GEMM, sizes read on device)
//   finish    deflated columns copied, new eigenvalues stored
// All data-dependent sizes (k = number of non-deflated poles) stay on the device; the host only
// enqueues, so the whole solve is asynchronous on one stream.
#include <vector>
#include "kernels.h"
#include "dc_core.cuh"

namespace zq {

struct MergeDesc { int off, n1, n2; };

struct DcLevel { int first, count, maxnm; bool uniform; };   // uniform: every merge of the level spans maxnm rows, block m at offset m * maxnm

struct DcWs {
  int n = 0;
  std::vector<DcLevel> levels;     // level 0 = leaves
  MergeDesc* merges = nullptr;     // device, all levels concatenated
  int* bounds = nullptr; int nbounds = 0;
  double *d = nullptr, *e = nullptr, *Q[2] = {nullptr, nullptr}, *S = nullptr;
  double *z = nullptr, *ds = nullptr, *zs = nullptr, *dlam = nullptr, *wz = nullptr, *z2 = nullptr;
  double *dfval = nullptr, *rcc = nullptr, *rss = nullptr, *zhat = nullptr, *lam = nullptr, *rho = nullptr;
  int *col = nullptr, *ndcol = nullptr, *dfcol = nullptr, *rc1 = nullptr, *rc2 = nullptr, *kcnt = nullptr, *perm = nullptr;
  int *ndtype = nullptr, *pinv = nullptr, *acol = nullptr;
  double* scale = nullptr;
  void* base = nullptr;
  long launches = 0;
};

namespace {

constexpr int GB = 64, GK = 16;
constexpr int KC = 8;   // ints per merge in kcnt: k, ndefl, nrot, k1, k2 (columns of type 1 / 2), spare

__global__ void __launch_bounds__(1024) k_dc_init(int n, const double* din, const double* ein, double* d, double* e,
                                                  double* scale) {
  __shared__ double sm[32];
  double m = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    m = fmax(m, fabs(din[i]));
    if (i + 1 < n) m = fmax(m, fabs(ein[i]));
  }
  m = WarpLanes::maxv(m);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
  __syncthreads();
  m = 0.0;
  for (int j = 0; j < (int)(blockDim.x >> 5); ++j) m = fmax(m, sm[j]);
  if (!(m > 0.0) || !isfinite(m)) m = 1.0;
  const double inv = 1.0 / m;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    d[i] = din[i] * inv;
    e[i] = (i + 1 < n) ? ein[i] * inv : 0.0;
  }
  if (threadIdx.x == 0) *scale = m;
}

__global__ void k_dc_tear(int nb, const int* bounds, double* d, const double* e) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nb) return;
  const int p = bounds[i];
  const double r = fabs(e[p - 1]);
  d[p - 1] -= r;
  d[p] -= r;
}

// one warp per leaf; 4 leaves per CTA
__global__ void __launch_bounds__(128) k_dc_leaf(int nleaf, const MergeDesc* leaves, double* d, const double* e,
                                                 double* Q, size_t ldq, int* info) {
  __shared__ double Zs[4][DC_LEAF * (DC_LEAF + 1)];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int li = blockIdx.x * 4 + w;
  if (li >= nleaf) return;
  const int off = leaves[li].off, m = leaves[li].n1;
  double dl[DC_LEAF], el[DC_LEAF];
#pragma unroll 1
  for (int i = 0; i < m; ++i) { dl[i] = d[off + i]; el[i] = (i + 1 < m) ? e[off + i] : 0.0; }
  double* Z = Zs[w];
  const int ldz = DC_LEAF + 1;
  for (int c = 0; c < m; ++c) Z[lane + c * ldz] = (lane == c) ? 1.0 : 0.0;
  __syncwarp();
  const int fail = leaf_ql<WarpLanes>(m, dl, el, Z, ldz);
  __syncwarp();
  if (lane < m) {
    for (int c = 0; c < m; ++c) Q[(size_t)(off + lane) + (size_t)(off + c) * ldq] = Z[lane + c * ldz];
  }
  if (lane == 0) {
    for (int i = 0; i < m; ++i) d[off + i] = dl[i];
    if (fail) atomicOr(info, 4);
  }
}

// z vector + rank sort of the children's eigenvalues
__global__ void __launch_bounds__(256) k_dc_prep(const MergeDesc* mg, const double* __restrict__ d,
                                                 const double* __restrict__ e, const double* __restrict__ Q, size_t ldq,
                                                 double* ds, double* zs, int* col, double* rho) {
  const MergeDesc md = mg[blockIdx.y];
  const int nm = md.n1 + md.n2, off = md.off;
  if ((int)(blockIdx.x * 256) >= nm) return;
  const int j = blockIdx.x * 256 + threadIdx.x;
  const double ecut = e[off + md.n1 - 1];
  __shared__ double dsm[256];
  const double dj = (j < nm) ? d[off + j] : 0.0;
  int rank = 0;
  for (int l0 = 0; l0 < nm; l0 += 256) {
    __syncthreads();
    if (l0 + (int)threadIdx.x < nm) dsm[threadIdx.x] = d[off + l0 + threadIdx.x];
    __syncthreads();
    const int lim = min(256, nm - l0);
    for (int l = 0; l < lim; ++l) {
      const double dv = dsm[l];
      rank += (dv < dj || (dv == dj && (l0 + l) < j)) ? 1 : 0;
    }
  }
  if (j < nm) {
    double zj;
    if (j < md.n1) zj = Q[(size_t)(off + md.n1 - 1) + (size_t)(off + j) * ldq];
    else zj = (ecut >= 0.0 ? 1.0 : -1.0) * Q[(size_t)(off + md.n1) + (size_t)(off + j) * ldq];
    zj *= 0.70710678118654752440;
    ds[off + rank] = dj;
    zs[off + rank] = zj;
    col[off + rank] = j;
  }
  if (j == 0) rho[blockIdx.y] = 2.0 * fabs(ecut);
}

__global__ void __launch_bounds__(32) k_dc_deflate(const MergeDesc* mg, const double* rho, const double* ds,
                                                   const double* zs, const int* col, double* dlam, double* wz,
                                                   double* z2, int* ndcol, int* ndtype, double* dfval, int* dfcol, int* rc1,
                                                   int* rc2, double* rcc, double* rss, int* pinv, int* acol, int* kcnt) {
  const MergeDesc md = mg[blockIdx.x];
  const int nm = md.n1 + md.n2, off = md.off, lane = threadIdx.x;
  DeflateOut o = deflate_scan<WarpLanes>(nm, md.n1, rho[blockIdx.x], ds + off, zs + off, col + off, dlam + off, wz + off,
                                         ndcol + off, ndtype + off, dfval + off, dfcol + off, rc1 + off, rc2 + off,
                                         rcc + off, rss + off);
  __syncwarp();
  __threadfence_block();
  for (int i = lane; i < o.k; i += 32) { const double v = wz[off + i]; z2[off + i] = v * v; }
  // rows of the secular eigenvector matrix S are stored grouped by column type [1 | 2 | 3] (stable), so
  // the merge GEMM multiplies the rows of child 1 with S-rows [0, k1+k2) and those of child 2 with [k1, k)
  int cnt[3] = {0, 0, 0};
  for (int l0 = 0; l0 < o.k; l0 += 32) {
    const int l = l0 + lane;
    const int t = (l < o.k) ? ndtype[off + l] : 0;
#pragma unroll
    for (int c = 0; c < 3; ++c) cnt[c] += __popc(__ballot_sync(0xffffffffu, t == c + 1));
  }
  int base[3] = {0, cnt[0], cnt[0] + cnt[1]};
  for (int l0 = 0; l0 < o.k; l0 += 32) {
    const int l = l0 + lane;
    const int t = (l < o.k) ? ndtype[off + l] : 0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const unsigned m = __ballot_sync(0xffffffffu, t == c + 1);
      if (t == c + 1) {
        const int pos = base[c] + __popc(m & ((1u << lane) - 1u));
        pinv[off + pos] = l;
        acol[off + pos] = ndcol[off + l];
      }
      base[c] += __popc(m);
    }
  }
  if (lane == 0) {
    kcnt[KC * blockIdx.x + 0] = o.k;
    kcnt[KC * blockIdx.x + 1] = o.ndefl;
    kcnt[KC * blockIdx.x + 2] = o.nrot;
    kcnt[KC * blockIdx.x + 3] = cnt[0];
    kcnt[KC * blockIdx.x + 4] = cnt[1];
  }
}

__global__ void __launch_bounds__(256) k_dc_rotate(const MergeDesc* mg, const int* kcnt, const int* rc1, const int* rc2,
                                                   const double* rcc, const double* rss, double* Q, size_t ldq) {
  const MergeDesc md = mg[blockIdx.y];
  const int nm = md.n1 + md.n2, off = md.off;
  const int nrot = kcnt[KC * blockIdx.y + 2];
  const int r = blockIdx.x * 256 + threadIdx.x;
  if (r >= nm || nrot == 0) return;
  for (int q = 0; q < nrot; ++q) {
    const int c1 = rc1[off + q], c2 = rc2[off + q];
    const double c = rcc[off + q], s = rss[off + q];
    double* px = Q + (size_t)(off + r) + (size_t)(off + c1) * ldq;
    double* py = Q + (size_t)(off + r) + (size_t)(off + c2) * ldq;
    const double x = *px, y = *py;
    *px = c * x + s * y;
    *py = c * y - s * x;
  }
}

__global__ void __launch_bounds__(256) k_dc_secular(const MergeDesc* mg, const int* kcnt, const double* rho,
                                                    const double* dlam, const double* z2, const int* __restrict__ pinv,
                                                    double* S, size_t lds, double* lam, int* info) {
  const MergeDesc md = mg[blockIdx.y];
  const int off = md.off;
  const int k = kcnt[KC * blockIdx.y];
  const int j = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (j >= k) return;
  const int lane = threadIdx.x & 31;
  int org;
  double mu;
  const int it = secular_root<WarpLanes>(j, k, dlam + off, z2 + off, rho[blockIdx.y], &org, &mu);
  const double dorg = dlam[off + org];
  double* Sc = S + (size_t)off + (size_t)(off + j) * lds;
  for (int r = lane; r < k; r += 32) Sc[r] = (dlam[off + pinv[off + r]] - dorg) - mu;   // S-row r <-> pole pinv[r]
  if (lane == 0) {
    lam[off + j] = dorg + mu;
    if (it >= 100) atomicOr(info, 2);
  }
}

__global__ void __launch_bounds__(256) k_dc_zhat(const MergeDesc* mg, const int* kcnt, const double* __restrict__ dlam,
                                                 const double* __restrict__ wz, const int* __restrict__ pinv,
                                                 const double* __restrict__ S, size_t lds, double* zhat) {
  const MergeDesc md = mg[blockIdx.y];
  const int off = md.off;
  const int k = kcnt[KC * blockIdx.y];
  const int r = blockIdx.x * 256 + threadIdx.x;      // S-row
  if (r >= k) return;
  const int i = pinv[off + r];                       // its pole
  const double di = dlam[off + i];
  const double* Sr = S + (size_t)(off + r) + (size_t)off * lds;
  double p = Sr[(size_t)i * lds];
  for (int j = 0; j < k; ++j) {
    if (j == i) continue;
    p *= Sr[(size_t)j * lds] / (di - dlam[off + j]);
  }
  zhat[off + r] = copysign(sqrt(fabs(p)), wz[off + i]);   // indexed by S-row
}

__global__ void __launch_bounds__(128) k_dc_vectors(const MergeDesc* mg, const int* kcnt, const double* __restrict__ zhat,
                                                    double* S, size_t lds) {
  const MergeDesc md = mg[blockIdx.y];
  const int off = md.off;
  const int k = kcnt[KC * blockIdx.y];
  const int j = blockIdx.x;
  if (j >= k) return;
  __shared__ double sm[32];
  double* Sc = S + (size_t)off + (size_t)(off + j) * lds;
  double ss = 0.0;
  for (int i = threadIdx.x; i < k; i += 128) {
    const double v = zhat[off + i] / Sc[i];
    Sc[i] = v;
    ss += v * v;
  }
  double v1[1] = {ss};
  block_sum<1>(v1, sm);
  const double inv = 1.0 / sqrt(v1[0]);
  for (int i = threadIdx.x; i < k; i += 128) Sc[i] *= inv;
}

// Merge GEMM on the FP64 tensor path (DMMA m8n8k4), sizes read on the device:
//   rows of child 1:  Qnew[off + r, off + j]      = sum_{p in [0, k1+k2)} Qold[off + r, off + acol[p]] * S[off + p, off + j]
//   rows of child 2:  Qnew[off + n1 + r, off + j] = sum_{p in [k1, k)}     Qold[off + n1 + r, off + acol[p]] * S[off + p, off + j]
// (the skipped products are exact zeros: a type-1 column vanishes in the rows of child 2 and vice versa).
// CTA tile 128 x 128, 8 warps (4 x 2) of 32 x 64, BK = 16, 3-stage cp.async (8-byte) pipeline.
constexpr int DG_BM = 128, DG_BN = 128, DG_BK = 16, DG_ST = 3, DG_LDA = DG_BM + 4, DG_LDB = DG_BK + 4;
constexpr int DG_STAGE = DG_BK * DG_LDA + DG_BN * DG_LDB;   // doubles per stage

ZQ_D void cp_async8(void* smem, const void* gmem, bool pred) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  const int sz = pred ? 8 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gmem), "r"(sz));
}

__global__ void __launch_bounds__(256) k_dc_gemm_mma(const MergeDesc* mg, const int* kcnt, const int* __restrict__ acol,
                                                     const double* __restrict__ Qold, const double* __restrict__ S,
                                                     double* __restrict__ Qnew, size_t ld, int cbeg, int cend) {
  const int merge = blockIdx.z >> 1, half = blockIdx.z & 1;
  const MergeDesc md = mg[merge];
  const int off = md.off;
  const int kfull = kcnt[KC * merge], k1 = kcnt[KC * merge + 3], k2 = kcnt[KC * merge + 4];
  const int k = kfull < cend ? kfull : cend;          // columns [cbeg, min(k, cend)) belong to this rank (multi-GPU top merge)
  const int rb = half ? md.n1 : 0, nr = half ? md.n2 : md.n1;
  const int kb = half ? k1 : 0, KK = half ? (kfull - k1) : (k1 + k2);
  const int r0 = blockIdx.x * DG_BM, c0 = cbeg + blockIdx.y * DG_BN;
  if (r0 >= nr || c0 >= k || KK <= 0) return;      // Qnew was zero-filled
  extern __shared__ __align__(16) unsigned char dg_smem_raw[];
  double* smem = reinterpret_cast<double*>(dg_smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = (warp & 3) * 32, wn = (warp >> 2) * 64;
  const int g = lane >> 2, q = lane & 3;
  const double* Abase = Qold + (size_t)(off + rb + r0) + (size_t)off * ld;
  const double* Bbase = S + (size_t)(off + kb) + (size_t)(off + c0) * ld;
  const int* ac = acol + off + kb;

  double acc[4][8][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  const int nk = (KK + DG_BK - 1) / DG_BK;
  auto issue = [&](int kt) {
    if (kt < nk) {
      double* sa = smem + (size_t)(kt % DG_ST) * DG_STAGE;
      double* sb = sa + DG_BK * DG_LDA;
      const int k0 = kt * DG_BK;
#pragma unroll
      for (int e = tid; e < DG_BM * DG_BK; e += 256) {
        const int m = e % DG_BM, kk = e / DG_BM;
        const bool ok = (r0 + m < nr) && (k0 + kk < KK);
        const double* src = ok ? Abase + m + (size_t)ac[k0 + kk] * ld : Qold;
        cp_async8(sa + kk * DG_LDA + m, src, ok);
      }
#pragma unroll
      for (int e = tid; e < DG_BN * DG_BK; e += 256) {
        const int kk = e % DG_BK, nn = e / DG_BK;
        const bool ok = (c0 + nn < k) && (k0 + kk < KK);
        const double* src = ok ? Bbase + (k0 + kk) + (size_t)nn * ld : S;
        cp_async8(sb + nn * DG_LDB + kk, src, ok);
      }
    }
    asm volatile("cp.async.commit_group;\n" ::);
  };
#pragma unroll
  for (int s2 = 0; s2 < DG_ST - 1; ++s2) issue(s2);
  for (int kt = 0; kt < nk; ++kt) {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(DG_ST - 2));
    __syncthreads();
    issue(kt + DG_ST - 1);
    const double* sa = smem + (size_t)(kt % DG_ST) * DG_STAGE;
    const double* sb = sa + DG_BK * DG_LDA;
#pragma unroll
    for (int k4 = 0; k4 < DG_BK; k4 += 4) {
      double a[4], b[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = sa[(k4 + q) * DG_LDA + wm + 8 * i + g];
#pragma unroll
      for (int j = 0; j < 8; ++j) b[j] = sb[(wn + 8 * j + g) * DG_LDB + k4 + q];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j)
          asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                       : "+d"(acc[i][j][0]), "+d"(acc[i][j][1]) : "d"(a[i]), "d"(b[j]));
    }
  }
  asm volatile("cp.async.wait_group 0;\n" ::);
  double* Cbase = Qnew + (size_t)(off + rb + r0) + (size_t)(off + c0) * ld;
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int c = wn + 8 * j + 2 * q + h;
      if (c0 + c >= k) continue;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = wm + 8 * i + g;
        if (r0 + r < nr) Cbase[(size_t)r + (size_t)c * ld] = acc[i][j][h];
      }
    }
}

__global__ void __launch_bounds__(256) k_dc_finish(const MergeDesc* mg, const int* kcnt, const int* __restrict__ dfcol,
                                                   const double* __restrict__ dfval, const double* __restrict__ lam,
                                                   const double* __restrict__ Qold, double* __restrict__ Qnew, size_t ld,
                                                   double* d, int cbeg, int cend) {
  const MergeDesc md = mg[blockIdx.z];
  const int nm = md.n1 + md.n2, off = md.off;
  const int k = kcnt[KC * blockIdx.z];
  const int c = blockIdx.y;            // column of the merged block
  const int r = blockIdx.x * 256 + threadIdx.x;
  if (c >= nm || (int)(blockIdx.x * 256) >= nm) return;
  if (c < k) {
    if (r == 0) d[off + c] = lam[off + c];
    return;
  }
  const int src = dfcol[off + c - k];
  if (r < nm && c >= cbeg && c < cend) Qnew[(size_t)(off + r) + (size_t)(off + c) * ld] = Qold[(size_t)(off + r) + (size_t)(off + src) * ld];
  if (r == 0) d[off + c] = dfval[off + c - k];
}

__global__ void __launch_bounds__(256) k_dc_final_sort(int n, const double* __restrict__ d, const double* scale,
                                                       double* wout, int* perm) {
  const int j = blockIdx.x * 256 + threadIdx.x;
  __shared__ double dsm[256];
  const double dj = (j < n) ? d[j] : 0.0;
  int rank = 0;
  for (int l0 = 0; l0 < n; l0 += 256) {
    __syncthreads();
    if (l0 + (int)threadIdx.x < n) dsm[threadIdx.x] = d[l0 + threadIdx.x];
    __syncthreads();
    const int lim = min(256, n - l0);
    for (int l = 0; l < lim; ++l) {
      const double dv = dsm[l];
      rank += (dv < dj || (dv == dj && (l0 + l) < j)) ? 1 : 0;
    }
  }
  if (j < n) {
    wout[rank] = dj * (*scale);
    perm[rank] = j;
  }
}

inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// ZQ_DC_SPLIT_LEVELS=0: only the top merge is split across the ranks (the round-1 behaviour); read at every solve
inline bool dc_split_levels() {
  const char* e = getenv("ZQ_DC_SPLIT_LEVELS");
  return !(e && atoi(e) == 0);
}

}  // namespace

static void build_tree(int n, std::vector<std::vector<MergeDesc>>& lv) {
  // segments per depth, root first
  std::vector<std::vector<std::pair<int, int>>> segs;
  segs.push_back({{0, n}});
  for (;;) {
    int mx = 0;
    for (auto& s : segs.back()) mx = s.second > mx ? s.second : mx;
    if (mx <= DC_LEAF) break;
    std::vector<std::pair<int, int>> nxt;
    for (auto& s : segs.back()) {
      const int h = s.second / 2;
      nxt.push_back({s.first, h});
      nxt.push_back({s.first + h, s.second - h});
    }
    segs.push_back(nxt);
  }
  const int depth = (int)segs.size();
  lv.clear();
  // level 0: leaves
  std::vector<MergeDesc> leaves;
  for (auto& s : segs[depth - 1]) leaves.push_back({s.first, s.second, 0});
  lv.push_back(leaves);
  for (int dd = depth - 2; dd >= 0; --dd) {
    std::vector<MergeDesc> ms;
    for (size_t i = 0; i < segs[dd].size(); ++i) {
      const int n1 = segs[dd + 1][2 * i].second, n2 = segs[dd + 1][2 * i + 1].second;
      ms.push_back({segs[dd][i].first, n1, n2});
    }
    lv.push_back(ms);
  }
}

size_t dc_bytes(int n) {
  const size_t nn = (size_t)n * n;
  return 3 * nn * sizeof(double) + (size_t)n * 32 * sizeof(double);
}

DcWs* dc_create(int n) {
  DcWs* ws = new DcWs();
  ws->n = n;
  std::vector<std::vector<MergeDesc>> lv;
  build_tree(n, lv);
  std::vector<MergeDesc> all;
  std::vector<int> bounds;
  for (auto& L : lv) {
    DcLevel dl;
    dl.first = (int)all.size();
    dl.count = (int)L.size();
    dl.maxnm = 0;
    for (auto& m : L) {
      all.push_back(m);
      const int nm = m.n1 + m.n2;
      dl.maxnm = nm > dl.maxnm ? nm : dl.maxnm;
    }
    dl.uniform = true;
    for (size_t i = 0; i < L.size(); ++i)
      if (L[i].n1 + L[i].n2 != dl.maxnm || L[i].off != (int)i * dl.maxnm) dl.uniform = false;
    ws->levels.push_back(dl);
  }
  for (auto& m : lv[0]) if (m.off > 0) bounds.push_back(m.off);
  ws->nbounds = (int)bounds.size();
  const size_t nn = (size_t)n * n;
  const size_t nd = (size_t)n;
  const int nmerge_max = (int)lv[0].size();
  // one slab
  size_t bytes = 0;
  auto take = [&](size_t b) { size_t o = bytes; bytes += (b + 255) & ~(size_t)255; return o; };
  const size_t oQ0 = take(nn * 8), oQ1 = take(nn * 8), oS = take(nn * 8);
  size_t od[14];
  for (int i = 0; i < 14; ++i) od[i] = take(nd * 8);
  size_t oi[9];
  for (int i = 0; i < 9; ++i) oi[i] = take(nd * 4);
  const size_t okc = take((size_t)nmerge_max * KC * 4), orho = take((size_t)nmerge_max * 8);
  const size_t omg = take(all.size() * sizeof(MergeDesc)), obd = take((bounds.size() + 1) * 4), osc = take(8);
  char* base = nullptr;
  if (cudaMalloc(&base, bytes) != cudaSuccess) { delete ws; return nullptr; }
  ws->base = base;
  ws->Q[0] = (double*)(base + oQ0); ws->Q[1] = (double*)(base + oQ1); ws->S = (double*)(base + oS);
  double** dp[14] = {&ws->d, &ws->e, &ws->z, &ws->ds, &ws->zs, &ws->dlam, &ws->wz, &ws->z2, &ws->dfval, &ws->rcc,
                     &ws->rss, &ws->zhat, &ws->lam, &ws->rho};
  for (int i = 0; i < 13; ++i) *dp[i] = (double*)(base + od[i]);
  ws->rho = (double*)(base + orho);
  int** ip[9] = {&ws->col, &ws->ndcol, &ws->dfcol, &ws->rc1, &ws->rc2, &ws->perm, &ws->ndtype, &ws->pinv, &ws->acol};
  for (int i = 0; i < 9; ++i) *ip[i] = (int*)(base + oi[i]);
  ws->kcnt = (int*)(base + okc);
  ws->merges = (MergeDesc*)(base + omg);
  ws->bounds = (int*)(base + obd);
  ws->scale = (double*)(base + osc);
  cudaMemcpy(ws->merges, all.data(), all.size() * sizeof(MergeDesc), cudaMemcpyHostToDevice);
  if (!bounds.empty()) cudaMemcpy(ws->bounds, bounds.data(), bounds.size() * 4, cudaMemcpyHostToDevice);
  return ws;
}

long dc_launches(const DcWs* ws) { return ws ? ws->launches : 0; }

void dc_destroy(DcWs* ws) {
  if (!ws) return;
  if (ws->base) cudaFree(ws->base);
  delete ws;
}

int dc_solve(DcWs* ws, int n, const double* d, const double* e, double* wout, double** Zres, int** perm, int* info,
             cudaStream_t st, const DcDist* dd) {
  const size_t ld = (size_t)n;
  static std::atomic<unsigned long long> attr_done{0};
  if (first_use_on_this_device(attr_done)) {
    cudaFuncSetAttribute(k_dc_gemm_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(DG_ST * DG_STAGE * sizeof(double)));
  }
  k_dc_init<<<1, 1024, 0, st>>>(n, d, e, ws->d, ws->e, ws->scale);
  if (ws->nbounds > 0) k_dc_tear<<<cdiv(ws->nbounds, 128), 128, 0, st>>>(ws->nbounds, ws->bounds, ws->d, ws->e);
  cudaMemsetAsync(ws->Q[0], 0, (size_t)n * n * sizeof(double), st);
  const DcLevel& L0 = ws->levels[0];
  k_dc_leaf<<<cdiv(L0.count, 4), 128, 0, st>>>(L0.count, ws->merges + L0.first, ws->d, ws->e, ws->Q[0], ld, info);
  int cur = 0;
  for (size_t li = 1; li < ws->levels.size(); ++li) {
    const DcLevel& L = ws->levels[li];
    const MergeDesc* mg = ws->merges + L.first;
    double* Qold = ws->Q[cur];
    double* Qnew = ws->Q[cur ^ 1];
    cudaMemsetAsync(Qnew, 0, (size_t)n * n * sizeof(double), st);
    k_dc_prep<<<dim3(cdiv(L.maxnm, 256), L.count), 256, 0, st>>>(mg, ws->d, ws->e, Qold, ld, ws->ds, ws->zs, ws->col, ws->rho);
    k_dc_deflate<<<L.count, 32, 0, st>>>(mg, ws->rho, ws->ds, ws->zs, ws->col, ws->dlam, ws->wz, ws->z2, ws->ndcol, ws->ndtype,
                                         ws->dfval, ws->dfcol, ws->rc1, ws->rc2, ws->rcc, ws->rss, ws->pinv, ws->acol, ws->kcnt);
    k_dc_rotate<<<dim3(cdiv(L.maxnm, 256), L.count), 256, 0, st>>>(mg, ws->kcnt, ws->rc1, ws->rc2, ws->rcc, ws->rss, Qold, ld);
    k_dc_secular<<<dim3(cdiv(L.maxnm, 8), L.count), 256, 0, st>>>(mg, ws->kcnt, ws->rho, ws->dlam, ws->z2, ws->pinv, ws->S, ld, ws->lam, info);
    k_dc_zhat<<<dim3(cdiv(L.maxnm, 256), L.count), 256, 0, st>>>(mg, ws->kcnt, ws->dlam, ws->wz, ws->pinv, ws->S, ld, ws->zhat);
    k_dc_vectors<<<dim3(L.maxnm, L.count), 128, 0, st>>>(mg, ws->kcnt, ws->zhat, ws->S, ld);
    {
      // multi-GPU: the merges of the top levels (the top merge, and below it the levels of 2 and 4 equal blocks while a block
      // still has >= 1024 columns and every rank >= 128 of them) are split by column blocks of the new eigenvector matrix --
      // rank g multiplies columns [g nm/G, (g+1) nm/G) of EVERY block of the level -- and all-gathered block by block (whole
      // columns of Qnew: zeros outside the diagonal block); the levels below are replicated
      const int nm = L.maxnm;
      const bool split = dd && dd->world > 1 && L.uniform && L.count <= 4 && (long long)nm * L.count == n && nm % dd->world == 0 &&
                         ((L.count == 1 && n >= 2048) || (L.count > 1 && nm >= 1024 && nm / dd->world >= 128 && dc_split_levels()));
      const int cbeg = split ? dd->rank * (nm / dd->world) : 0;
      const int cend = split ? cbeg + nm / dd->world : 0x7fffffff;
      const int ncolsg = split ? nm / dd->world : L.maxnm;
      const int n1max = (L.maxnm + 1) / 2;
      k_dc_gemm_mma<<<dim3(cdiv(n1max, DG_BM), cdiv(ncolsg, DG_BN), 2 * L.count), 256, DG_ST * DG_STAGE * sizeof(double), st>>>(
          mg, ws->kcnt, ws->acol, Qold, ws->S, Qnew, ld, cbeg, cend);
      k_dc_finish<<<dim3(cdiv(L.maxnm, 256), L.maxnm, L.count), 256, 0, st>>>(mg, ws->kcnt, ws->dfcol, ws->dfval, ws->lam, Qold, Qnew,
                                                                            ld, ws->d, cbeg, cend);
      if (split) {
        for (int m = 0; m < L.count; ++m) {
          const int rc = dd->allgather(Qnew + (size_t)m * nm * ld, (size_t)(nm / dd->world) * n, dd->rank, st);
          if (rc) return rc;
        }
      }
    }
    cur ^= 1;
  }
  k_dc_final_sort<<<cdiv(n, 256), 256, 0, st>>>(n, ws->d, ws->scale, wout, ws->perm);
  ws->launches = 4 + 9 * (long)(ws->levels.size() - 1) + 1;
  *Zres = ws->Q[cur];
  *perm = ws->perm;
  cudaError_t err = cudaGetLastError();
  return err == cudaSuccess ? 0 : -(int)err;
}

}  // namespace zq
