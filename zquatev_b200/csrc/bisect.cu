// K9: eigenvalues only (jobz = 'N') of the real symmetric tridiagonal by Sturm-sequence
// bisection, one thread per eigenvalue index.  Used by the values-only entry point
// (BASELINE config 4); the reference has no such mode -- zhbev is always called with "V"
// (zquatev.cc:84).
#include "kernels.h"
#include <float.h>

namespace zq {
namespace {

__global__ void __launch_bounds__(1024) k_gersh(int n, const double* d, const double* e, double* out) {
  __shared__ double slo[32], shi[32];
  double lo = DBL_MAX, hi = -DBL_MAX, emax = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double a = (i > 0) ? fabs(e[i - 1]) : 0.0;
    const double b = (i + 1 < n) ? fabs(e[i]) : 0.0;
    lo = fmin(lo, d[i] - a - b);
    hi = fmax(hi, d[i] + a + b);
    emax = fmax(emax, b);
  }
  for (int o = 16; o > 0; o >>= 1) {
    lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) { slo[threadIdx.x >> 5] = lo; shi[threadIdx.x >> 5] = hi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int j = 1; j < (int)(blockDim.x >> 5); ++j) { lo = fmin(lo, slo[j]); hi = fmax(hi, shi[j]); }
    const double nrm = fmax(fabs(lo), fabs(hi));
    out[0] = lo - 2.0 * DBL_EPSILON * nrm * n - 2.0 * DBL_MIN;
    out[1] = hi + 2.0 * DBL_EPSILON * nrm * n + 2.0 * DBL_MIN;
    out[2] = nrm;
  }
}

// number of eigenvalues < x
ZQ_D int sturm(int n, const double* __restrict__ d, const double* __restrict__ e2, double x, double pivmin) {
  int cnt = 0;
  double q = d[0] - x;
  if (fabs(q) < pivmin) q = -pivmin;
  cnt += (q < 0.0);
  for (int i = 1; i < n; ++i) {
    q = d[i] - x - e2[i - 1] / q;
    if (fabs(q) < pivmin) q = -pivmin;
    cnt += (q < 0.0);
  }
  return cnt;
}

__global__ void __launch_bounds__(128) k_bisect(int n, const double* __restrict__ d, const double* __restrict__ e2,
                                                const double* __restrict__ bnd, double* w, int jlo, int jhi) {
  const int j = jlo + blockIdx.x * 128 + threadIdx.x;
  if (j >= jhi) return;
  double lo = bnd[0], hi = bnd[1];
  const double nrm = bnd[2];
  const double pivmin = DBL_MIN * fmax(1.0, nrm * nrm);
  const double atol = 2.0 * DBL_EPSILON * nrm + 2.0 * pivmin;
  for (int it = 0; it < 120; ++it) {
    const double mid = 0.5 * (lo + hi);
    if (hi - lo <= atol || mid <= lo || mid >= hi) break;
    if (sturm(n, d, e2, mid, pivmin) > j) hi = mid; else lo = mid;
  }
  w[j] = 0.5 * (lo + hi);
}

__global__ void k_sq(int n, const double* e, double* e2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) e2[i] = e[i] * e[i];
}

}  // namespace

// scratch: needs n + 3 doubles.  Eigenvalue indices [jlo, jhi) only (multi-GPU: one index range per rank, SURVEY.md 8e;
// every index is an independent bisection, so the ranges need no exchange until the final all-gather).
void launch_bisect(int n, const double* d, const double* e, double* w, double* scratch, cudaStream_t st, int jlo, int jhi) {
  double* e2 = scratch;
  double* bnd = scratch + n;
  if (jhi < 0 || jhi > n) jhi = n;
  if (jlo < 0) jlo = 0;
  k_sq<<<(n + 255) / 256, 256, 0, st>>>(n, e, e2);
  k_gersh<<<1, 1024, 0, st>>>(n, d, e, bnd);
  if (jhi > jlo) k_bisect<<<(jhi - jlo + 127) / 128, 128, 0, st>>>(n, d, e2, bnd, w, jlo, jhi);
}

}  // namespace zq
