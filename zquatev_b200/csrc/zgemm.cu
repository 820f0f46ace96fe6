// K4 / K6: FP64 complex GEMM  C = alpha * op(A) * op(B) + beta * C  (column-major).
//
// Serves (a) the trailing rank-2k update of the tridiagonalisation -- the ONE stacked
// contraction [D;E] -= L R^H that replaces the eight zgemm3m calls of the reference
// (blocked.cc:424-474) -- with `lower` restricting work to the lower triangles, and (b) the
// compact-WY back-transformation Y = Phi(V)^H X, X -= Phi(V) (T Y) that replaces the
// reference's forward accumulation of Q (blocked.cc:477-544) and final zgemm3m pair
// (zquatev.cc:87-90).
//
// v1 kernel: 64x64x16 shared-memory tiles, 256 threads, 4x4 complex register tile per thread
// on the FP64 FMA pipe, register-prefetch double buffering.
#include "kernels.h"

namespace zq {
namespace {

constexpr int BM = 64, BN = 64, BK = 16, LDS = BM + 1;

template <int TA, int TB>
__global__ void __launch_bounds__(256)
k_zgemm(int M, int N, int K, cplx alpha, const cplx* __restrict__ A, size_t lda, const cplx* __restrict__ B,
        size_t ldb, cplx beta, cplx* __restrict__ C, size_t ldc, int lower, size_t sA, size_t sB, size_t sC) {
  const int bx = blockIdx.x, by = blockIdx.y;
  if (lower && bx < by) return;
  A += (size_t)blockIdx.z * sA;
  B += (size_t)blockIdx.z * sB;
  C += (size_t)blockIdx.z * sC;
  __shared__ cplx As[BK][LDS];
  __shared__ cplx Bs[BK][LDS];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int r0 = bx * BM, c0 = by * BN;

  cplx acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = cmake(0, 0);

  cplx ra[4], rb[4];
  auto gload = [&](int k0) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (TA == 0) {  // A is M x K: contiguous along rows
        const int r = r0 + (tid & 63), kk = k0 + (tid >> 6) + 4 * j;
        ra[j] = (r < M && kk < K) ? A[(size_t)r + (size_t)kk * lda] : cmake(0, 0);
      } else {        // A is K x M, op = conj transpose: contiguous along k
        const int kk = k0 + (tid & 15), r = r0 + (tid >> 4) + 16 * j;
        ra[j] = (r < M && kk < K) ? cconj(A[(size_t)kk + (size_t)r * lda]) : cmake(0, 0);
      }
      if (TB == 0) {  // B is K x N: contiguous along k
        const int kk = k0 + (tid & 15), c = c0 + (tid >> 4) + 16 * j;
        rb[j] = (c < N && kk < K) ? B[(size_t)kk + (size_t)c * ldb] : cmake(0, 0);
      } else {        // B is N x K, op = conj transpose: contiguous along columns of C
        const int c = c0 + (tid & 63), kk = k0 + (tid >> 6) + 4 * j;
        rb[j] = (c < N && kk < K) ? cconj(B[(size_t)c + (size_t)kk * ldb]) : cmake(0, 0);
      }
    }
  };
  auto sstore = [&]() {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (TA == 0) As[(tid >> 6) + 4 * j][tid & 63] = ra[j];
      else         As[tid & 15][(tid >> 4) + 16 * j] = ra[j];
      if (TB == 0) Bs[tid & 15][(tid >> 4) + 16 * j] = rb[j];
      else         Bs[(tid >> 6) + 4 * j][tid & 63] = rb[j];
    }
  };

  gload(0);
  for (int k0 = 0; k0 < K; k0 += BK) {
    sstore();
    __syncthreads();
    if (k0 + BK < K) gload(k0 + BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      cplx a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][tx + 16 * i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][ty + 16 * j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) cfma(acc[i][j], a[i], b[j]);
    }
    __syncthreads();
  }
  const bool bzero = (beta.x == 0.0 && beta.y == 0.0);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = c0 + ty + 16 * j;
    if (c >= N) continue;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = r0 + tx + 16 * i;
      if (r >= M || (lower && r < c)) continue;
      cplx v = cmul(alpha, acc[i][j]);
      cplx* cp = C + (size_t)r + (size_t)c * ldc;
      if (!bzero) { cplx o = *cp; cfma(v, beta, o); }
      *cp = v;
    }
  }
}

}  // namespace

void launch_zgemm(int ta, int tb, int M, int N, int K, cplx alpha, const cplx* A, size_t lda, const cplx* B,
                  size_t ldb, cplx beta, cplx* C, size_t ldc, int lower, int batch, size_t sA, size_t sB,
                  size_t sC, cudaStream_t st) {
  if (M <= 0 || N <= 0 || batch <= 0) return;
  dim3 g((M + BM - 1) / BM, (N + BN - 1) / BN, batch);
  if (ta == 0 && tb == 0) k_zgemm<0, 0><<<g, 256, 0, st>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, lower, sA, sB, sC);
  else if (ta == 0 && tb == 1) k_zgemm<0, 1><<<g, 256, 0, st>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, lower, sA, sB, sC);
  else if (ta == 1 && tb == 0) k_zgemm<1, 0><<<g, 256, 0, st>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, lower, sA, sB, sC);
  else k_zgemm<1, 1><<<g, 256, 0, st>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, lower, sA, sB, sC);
}

}  // namespace zq
