// Input scaling (the zlascl step of LAPACK's zheev driver, which the reference reaches through zheev / zhbev in
// test.cc:84-95 and zquatev.cc:84): when the largest entry of the matrix lies outside [rmin, rmax] =
// [sqrt(safmin/eps), sqrt(eps/safmin)] the matrix is scaled into that range before the reduction and the eigenvalues
// are scaled back afterwards.  Inside the range the plain sums of squares of the reflector kernels (panel.cu,
// small.cu) can neither overflow nor lose the norm to underflow, which is what LAPACK's scaled dznrm2 / zlarfg
// protect against (reference blocked.cc:176,345).  Only the lower triangles of D and E are looked at and scaled --
// nothing else is ever read.  Eigenvectors are invariant under the scaling.
#include "kernels.h"

namespace zq {
namespace {

constexpr int SC_NT = 256, SC_COLS = 8;

// max |re|, |im| over the lower triangles: grid (row blocks of 1024, column groups of SC_COLS)
__global__ void __launch_bounds__(SC_NT) k_amax_lower(const cplx* __restrict__ A, size_t lda, int n, double* __restrict__ part) {
  __shared__ double sm[32];
  const int c0 = blockIdx.y * SC_COLS;
  double mx = 0.0;
  for (int c = c0; c < min(n, c0 + SC_COLS); ++c) {
    for (int r = blockIdx.x * (4 * SC_NT) + threadIdx.x; r < min(n, (int)(blockIdx.x + 1) * (4 * SC_NT)); r += SC_NT) {
      if (r < c) continue;
      const cplx d = A[(size_t)r + (size_t)c * lda];
      // diagonal of D: only the real part is referenced (zlanhe convention).  fmax ignores NaN: a NaN input surfaces
      // later as info > 0
      mx = fmax(mx, r == c ? fabs(d.x) : fmax(fabs(d.x), fabs(d.y)));
      if (r > c) {
        const cplx e = A[(size_t)(n + r) + (size_t)c * lda];
        mx = fmax(mx, fmax(fabs(e.x), fabs(e.y)));
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < SC_NT / 32; ++w) mx = fmax(mx, sm[w]);
    part[blockIdx.x + gridDim.x * blockIdx.y] = mx;
  }
}

// sc[0] = sigma applied to the matrix (1 = none), sc[1] = 1/sigma applied to the eigenvalues
__global__ void __launch_bounds__(SC_NT) k_scale_decide(const double* __restrict__ part, int np, double* __restrict__ sc) {
  __shared__ double sm[32];
  double mx = 0.0;
  for (int i = threadIdx.x; i < np; i += SC_NT) mx = fmax(mx, part[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < SC_NT / 32; ++w) mx = fmax(mx, sm[w]);
    const double safmin = 2.2250738585072014e-308, eps = 1.1102230246251565e-16;
    const double rmin = sqrt(safmin / eps), rmax = sqrt(eps / safmin);
    double sigma = 1.0;
    if (mx > 0.0 && isfinite(mx)) {
      if (mx < rmin) sigma = rmin / mx;
      else if (mx > rmax) sigma = rmax / mx;
    }
    sc[0] = sigma;
    sc[1] = 1.0 / sigma;
  }
}

__global__ void __launch_bounds__(SC_NT) k_scale_lower(cplx* __restrict__ A, size_t lda, int n, const double* __restrict__ sc) {
  const double sigma = sc[0];
  if (sigma == 1.0) return;                              // the common case: nothing to do
  const int c0 = blockIdx.y * SC_COLS;
  for (int c = c0; c < min(n, c0 + SC_COLS); ++c) {
    for (int r = blockIdx.x * (4 * SC_NT) + threadIdx.x; r < min(n, (int)(blockIdx.x + 1) * (4 * SC_NT)); r += SC_NT) {
      if (r < c) continue;
      cplx* d = A + (size_t)r + (size_t)c * lda;
      *d = cscale(*d, sigma);
      if (r > c) {
        cplx* e = A + (size_t)(n + r) + (size_t)c * lda;
        *e = cscale(*e, sigma);
      }
    }
  }
}

__global__ void __launch_bounds__(SC_NT) k_unscale_eig(int n, double* __restrict__ eig, const double* __restrict__ sc) {
  const double inv = sc[1];
  if (inv == 1.0) return;
  const int i = blockIdx.x * SC_NT + threadIdx.x;
  if (i < n) eig[i] *= inv;
}

}  // namespace

size_t scale_scratch_doubles(int n) {
  const size_t gx = (size_t)(n + 4 * SC_NT - 1) / (4 * SC_NT), gy = (size_t)(n + SC_COLS - 1) / SC_COLS;
  return gx * gy + 8;
}

// scratch: scale_scratch_doubles(n) doubles; scratch[0..1] = (sigma, 1/sigma) afterwards
void launch_scale_input(cplx* A, size_t lda, int n, double* scratch, cudaStream_t st) {
  const dim3 g((n + 4 * SC_NT - 1) / (4 * SC_NT), (n + SC_COLS - 1) / SC_COLS);
  k_amax_lower<<<g, SC_NT, 0, st>>>(A, lda, n, scratch + 8);
  k_scale_decide<<<1, SC_NT, 0, st>>>(scratch + 8, (int)(g.x * g.y), scratch);
  k_scale_lower<<<g, SC_NT, 0, st>>>(A, lda, n, scratch);
}

void launch_unscale_eig(int n, double* eig, const double* scratch, cudaStream_t st) {
  k_unscale_eig<<<(n + SC_NT - 1) / SC_NT, SC_NT, 0, st>>>(n, eig, scratch);
}

}  // namespace zq
