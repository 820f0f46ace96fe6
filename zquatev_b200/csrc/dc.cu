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

struct DcLevel { int first, count, maxnm; };

struct DcWs {
  int n = 0;
  std::vector<DcLevel> levels;     // level 0 = leaves
  MergeDesc* merges = nullptr;     // device, all levels concatenated
  int* bounds = nullptr; int nbounds = 0;
  double *d = nullptr, *e = nullptr, *Q[2] = {nullptr, nullptr}, *S = nullptr;
  double *z = nullptr, *ds = nullptr, *zs = nullptr, *dlam = nullptr, *wz = nullptr, *z2 = nullptr;
  double *dfval = nullptr, *rcc = nullptr, *rss = nullptr, *zhat = nullptr, *lam = nullptr, *rho = nullptr;
  int *col = nullptr, *ndcol = nullptr, *dfcol = nullptr, *rc1 = nullptr, *rc2 = nullptr, *kcnt = nullptr, *perm = nullptr;
  double* scale = nullptr;
  void* base = nullptr;
  long launches = 0;
};

namespace {

constexpr int GB = 64, GK = 16;

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
                                                   double* z2, int* ndcol, double* dfval, int* dfcol, int* rc1,
                                                   int* rc2, double* rcc, double* rss, int* kcnt) {
  const MergeDesc md = mg[blockIdx.x];
  const int nm = md.n1 + md.n2, off = md.off;
  DeflateOut o = deflate_scan<WarpLanes>(nm, rho[blockIdx.x], ds + off, zs + off, col + off, dlam + off, wz + off,
                                         ndcol + off, dfval + off, dfcol + off, rc1 + off, rc2 + off, rcc + off,
                                         rss + off);
  __syncwarp();
  __threadfence_block();
  for (int i = threadIdx.x; i < o.k; i += 32) { const double v = wz[off + i]; z2[off + i] = v * v; }
  if (threadIdx.x == 0) {
    kcnt[3 * blockIdx.x + 0] = o.k;
    kcnt[3 * blockIdx.x + 1] = o.ndefl;
    kcnt[3 * blockIdx.x + 2] = o.nrot;
  }
}

__global__ void __launch_bounds__(256) k_dc_rotate(const MergeDesc* mg, const int* kcnt, const int* rc1, const int* rc2,
                                                   const double* rcc, const double* rss, double* Q, size_t ldq) {
  const MergeDesc md = mg[blockIdx.y];
  const int nm = md.n1 + md.n2, off = md.off;
  const int nrot = kcnt[3 * blockIdx.y + 2];
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
                                                    const double* dlam, const double* z2, double* S, size_t lds,
                                                    double* lam, int* info) {
  const MergeDesc md = mg[blockIdx.y];
  const int off = md.off;
  const int k = kcnt[3 * blockIdx.y];
  const int j = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (j >= k) return;
  const int lane = threadIdx.x & 31;
  int org;
  double mu;
  const int it = secular_root<WarpLanes>(j, k, dlam + off, z2 + off, rho[blockIdx.y], &org, &mu);
  const double dorg = dlam[off + org];
  double* Sc = S + (size_t)off + (size_t)(off + j) * lds;
  for (int i = lane; i < k; i += 32) Sc[i] = (dlam[off + i] - dorg) - mu;
  if (lane == 0) {
    lam[off + j] = dorg + mu;
    if (it >= 100) atomicOr(info, 2);
  }
}

__global__ void __launch_bounds__(256) k_dc_zhat(const MergeDesc* mg, const int* kcnt, const double* __restrict__ dlam,
                                                 const double* __restrict__ wz, const double* __restrict__ S, size_t lds,
                                                 double* zhat) {
  const MergeDesc md = mg[blockIdx.y];
  const int off = md.off;
  const int k = kcnt[3 * blockIdx.y];
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= k) return;
  const double di = dlam[off + i];
  const double* Sr = S + (size_t)(off + i) + (size_t)off * lds;
  double p = Sr[(size_t)i * lds];
  for (int j = 0; j < k; ++j) {
    if (j == i) continue;
    p *= Sr[(size_t)j * lds] / (di - dlam[off + j]);
  }
  zhat[off + i] = copysign(sqrt(fabs(p)), wz[off + i]);
}

__global__ void __launch_bounds__(128) k_dc_vectors(const MergeDesc* mg, const int* kcnt, const double* __restrict__ zhat,
                                                    double* S, size_t lds) {
  const MergeDesc md = mg[blockIdx.y];
  const int off = md.off;
  const int k = kcnt[3 * blockIdx.y];
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

// Qnew[off + r, off + j] = sum_l Qold[off + r, off + ndcol[l]] * S[off + l, off + j]   (r < nm; j, l < k)
__global__ void __launch_bounds__(256) k_dc_gemm(const MergeDesc* mg, const int* kcnt, const int* __restrict__ ndcol,
                                                 const double* __restrict__ Qold, const double* __restrict__ S,
                                                 double* __restrict__ Qnew, size_t ld) {
  const MergeDesc md = mg[blockIdx.z];
  const int nm = md.n1 + md.n2, off = md.off;
  const int k = kcnt[3 * blockIdx.z];
  const int r0 = blockIdx.x * GB, c0 = blockIdx.y * GB;
  if (r0 >= nm || c0 >= k) return;
  __shared__ double As[GK][GB + 1];
  __shared__ double Bs[GK][GB + 1];
  __shared__ int cidx[GK];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
  for (int k0 = 0; k0 < k; k0 += GK) {
    __syncthreads();
    if (tid < GK) cidx[tid] = (k0 + tid < k) ? ndcol[off + k0 + tid] : 0;
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      {  // A: contiguous along rows
        const int r = r0 + (tid & 63), kk = (tid >> 6) + 4 * q;
        double v = 0.0;
        if (r < nm && k0 + kk < k) v = Qold[(size_t)(off + r) + (size_t)(off + cidx[kk]) * ld];
        As[kk][tid & 63] = v;
      }
      {  // B: contiguous along l
        const int kk = tid & 15, c = c0 + (tid >> 4) + 16 * q;
        double v = 0.0;
        if (c < k && k0 + kk < k) v = S[(size_t)(off + k0 + kk) + (size_t)(off + c) * ld];
        Bs[kk][(tid >> 4) + 16 * q] = v;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < GK; ++kk) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][tx + 16 * i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][ty + 16 * j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = c0 + ty + 16 * j;
    if (c >= k) continue;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = r0 + tx + 16 * i;
      if (r < nm) Qnew[(size_t)(off + r) + (size_t)(off + c) * ld] = acc[i][j];
    }
  }
}

__global__ void __launch_bounds__(256) k_dc_finish(const MergeDesc* mg, const int* kcnt, const int* __restrict__ dfcol,
                                                   const double* __restrict__ dfval, const double* __restrict__ lam,
                                                   const double* __restrict__ Qold, double* __restrict__ Qnew, size_t ld,
                                                   double* d) {
  const MergeDesc md = mg[blockIdx.z];
  const int nm = md.n1 + md.n2, off = md.off;
  const int k = kcnt[3 * blockIdx.z];
  const int c = blockIdx.y;            // column of the merged block
  const int r = blockIdx.x * 256 + threadIdx.x;
  if (c >= nm || (int)(blockIdx.x * 256) >= nm) return;
  if (c < k) {
    if (r == 0) d[off + c] = lam[off + c];
    return;
  }
  const int src = dfcol[off + c - k];
  if (r < nm) Qnew[(size_t)(off + r) + (size_t)(off + c) * ld] = Qold[(size_t)(off + r) + (size_t)(off + src) * ld];
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
  size_t oi[6];
  for (int i = 0; i < 6; ++i) oi[i] = take(nd * 4);
  const size_t okc = take((size_t)nmerge_max * 3 * 4), orho = take((size_t)nmerge_max * 8);
  const size_t omg = take(all.size() * sizeof(MergeDesc)), obd = take((bounds.size() + 1) * 4), osc = take(8);
  char* base = nullptr;
  if (cudaMalloc(&base, bytes) != cudaSuccess) { delete ws; return nullptr; }
  ws->base = base;
  ws->Q[0] = (double*)(base + oQ0); ws->Q[1] = (double*)(base + oQ1); ws->S = (double*)(base + oS);
  double** dp[14] = {&ws->d, &ws->e, &ws->z, &ws->ds, &ws->zs, &ws->dlam, &ws->wz, &ws->z2, &ws->dfval, &ws->rcc,
                     &ws->rss, &ws->zhat, &ws->lam, &ws->rho};
  for (int i = 0; i < 13; ++i) *dp[i] = (double*)(base + od[i]);
  ws->rho = (double*)(base + orho);
  int** ip[6] = {&ws->col, &ws->ndcol, &ws->dfcol, &ws->rc1, &ws->rc2, &ws->perm};
  for (int i = 0; i < 6; ++i) *ip[i] = (int*)(base + oi[i]);
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
             cudaStream_t st) {
  const size_t ld = (size_t)n;
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
    k_dc_deflate<<<L.count, 32, 0, st>>>(mg, ws->rho, ws->ds, ws->zs, ws->col, ws->dlam, ws->wz, ws->z2, ws->ndcol,
                                         ws->dfval, ws->dfcol, ws->rc1, ws->rc2, ws->rcc, ws->rss, ws->kcnt);
    k_dc_rotate<<<dim3(cdiv(L.maxnm, 256), L.count), 256, 0, st>>>(mg, ws->kcnt, ws->rc1, ws->rc2, ws->rcc, ws->rss, Qold, ld);
    k_dc_secular<<<dim3(cdiv(L.maxnm, 8), L.count), 256, 0, st>>>(mg, ws->kcnt, ws->rho, ws->dlam, ws->z2, ws->S, ld, ws->lam, info);
    k_dc_zhat<<<dim3(cdiv(L.maxnm, 256), L.count), 256, 0, st>>>(mg, ws->kcnt, ws->dlam, ws->wz, ws->S, ld, ws->zhat);
    k_dc_vectors<<<dim3(L.maxnm, L.count), 128, 0, st>>>(mg, ws->kcnt, ws->zhat, ws->S, ld);
    k_dc_gemm<<<dim3(cdiv(L.maxnm, GB), cdiv(L.maxnm, GB), L.count), 256, 0, st>>>(mg, ws->kcnt, ws->ndcol, Qold, ws->S, Qnew, ld);
    k_dc_finish<<<dim3(cdiv(L.maxnm, 256), L.maxnm, L.count), 256, 0, st>>>(mg, ws->kcnt, ws->dfcol, ws->dfval, ws->lam, Qold, Qnew, ld, ws->d);
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
