// Shared device helpers for the zquatev B200 kernels (sm_100a).
//
// Conventions: a quaternion q = a + j b is a pair of complex numbers;
// complex numbers are double2 (x = re, y = im); all matrices are column-major.
// The structured matrix of the reference, zquatev.h:40-44,
//     M = [[D, -conj(E)], [E, conj(D)]],  D = D^H, E = -E^T,
// is held as ONE device array `A` of 2n rows and n columns with leading dimension
// lda (>= 2n): D(r,c) = A[r + c*lda], E(r,c) = A[n + r + c*lda] -- exactly the left
// half of the caller's array, so no repack (zquatev.cc:48-54) is needed.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

typedef double2 cplx;

#define ZQ_HD __host__ __device__ __forceinline__
#define ZQ_D __device__ __forceinline__

ZQ_HD cplx cmake(double r, double i) { cplx z; z.x = r; z.y = i; return z; }
ZQ_HD cplx cconj(cplx a) { return cmake(a.x, -a.y); }
ZQ_HD cplx cneg(cplx a) { return cmake(-a.x, -a.y); }
ZQ_HD cplx cadd(cplx a, cplx b) { return cmake(a.x + b.x, a.y + b.y); }
ZQ_HD cplx csub(cplx a, cplx b) { return cmake(a.x - b.x, a.y - b.y); }
ZQ_HD cplx cmul(cplx a, cplx b) { return cmake(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
ZQ_HD cplx cscale(cplx a, double s) { return cmake(a.x * s, a.y * s); }
// acc += a*b
ZQ_HD void cfma(cplx& acc, cplx a, cplx b) {
  acc.x = fma(a.x, b.x, acc.x); acc.x = fma(-a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y); acc.y = fma(a.y, b.x, acc.y);
}
// acc += conj(a)*b
ZQ_HD void cfma_ca(cplx& acc, cplx a, cplx b) {
  acc.x = fma(a.x, b.x, acc.x); acc.x = fma(a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y); acc.y = fma(-a.y, b.x, acc.y);
}
// acc -= a*b
ZQ_HD void cfms(cplx& acc, cplx a, cplx b) {
  acc.x = fma(-a.x, b.x, acc.x); acc.x = fma(a.y, b.y, acc.x);
  acc.y = fma(-a.x, b.y, acc.y); acc.y = fma(-a.y, b.x, acc.y);
}
// acc -= conj(a)*b
ZQ_HD void cfms_ca(cplx& acc, cplx a, cplx b) {
  acc.x = fma(-a.x, b.x, acc.x); acc.x = fma(-a.y, b.y, acc.x);
  acc.y = fma(-a.x, b.y, acc.y); acc.y = fma(a.y, b.x, acc.y);
}

struct quat { cplx a, b; };
ZQ_HD quat qmake(cplx a, cplx b) { quat q; q.a = a; q.b = b; return q; }
ZQ_HD quat qzero() { return qmake(cmake(0, 0), cmake(0, 0)); }
// (a + jb)(c + jd) = (ac - conj(b) d) + j (bc + conj(a) d)
ZQ_HD quat qmul(quat p, quat q) {
  quat r;
  r.a = cmul(p.a, q.a); cfms_ca(r.a, p.b, q.b);
  r.b = cmul(p.b, q.a); cfma_ca(r.b, p.a, q.b);
  return r;
}
// acc += p*q
ZQ_HD void qfma(quat& acc, quat p, quat q) {
  cfma(acc.a, p.a, q.a); cfms_ca(acc.a, p.b, q.b);
  cfma(acc.b, p.b, q.a); cfma_ca(acc.b, p.a, q.b);
}
// acc -= p*q
ZQ_HD void qfms(quat& acc, quat p, quat q) {
  cfms(acc.a, p.a, q.a); cfma_ca(acc.a, p.b, q.b);
  cfms(acc.b, p.b, q.a); cfms_ca(acc.b, p.a, q.b);
}
// acc += p^* q  (quaternion inner-product term: p^* = conj(a) - j b)
//   (conj(pa) - j pb)(qa + j qb) = conj(pa) qa + conj(pb) qb + j (pa qb - pb qa)
ZQ_HD void qfma_cj(quat& acc, quat p, quat q) {
  cfma_ca(acc.a, p.a, q.a); cfma_ca(acc.a, p.b, q.b);
  cfma(acc.b, p.a, q.b); cfms(acc.b, p.b, q.a);
}
ZQ_HD quat qconj(quat p) { return qmake(cconj(p.a), cneg(p.b)); }
ZQ_HD quat qscale(quat p, double s) { return qmake(cscale(p.a, s), cscale(p.b, s)); }
ZQ_HD quat qadd(quat p, quat q) { return qmake(cadd(p.a, q.a), cadd(p.b, q.b)); }
ZQ_HD double qnorm2(quat p) { return p.a.x * p.a.x + p.a.y * p.a.y + p.b.x * p.b.x + p.b.y * p.b.y; }

#ifdef __CUDACC__
ZQ_D double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
ZQ_D cplx warp_sum(cplx v) { v.x = warp_sum(v.x); v.y = warp_sum(v.y); return v; }
ZQ_D quat warp_sum(quat v) { v.a = warp_sum(v.a); v.b = warp_sum(v.b); return v; }

// Programmatic dependent launch (PDL): the per-column kernels of the reduction are short and strictly dependent, so
// their launch latency is on the critical path n times.  Each chain kernel starts with pdl_enter(): it lets the NEXT
// kernel of the stream be scheduled as soon as every CTA of this one is resident (its CTAs then sit in
// griddepcontrol.wait) and itself waits until the PREVIOUS grid has completed and its writes are visible -- so
// nothing that reads or writes global memory may precede it.  Both instructions are no-ops in a launch without the
// programmatic-serialization attribute (launch_chain, kernels.h).
ZQ_D void pdl_enter() {
  asm volatile("griddepcontrol.launch_dependents;");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

// streaming 16-byte load that does not allocate in L1 (matrix data is touched once per pass)
ZQ_D cplx ld_stream(const cplx* p) {
  cplx v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}

// Deterministic block-wide sum of NV doubles (blockDim.x multiple of 32, <= 1024).
// Result valid in ALL threads.  `sm` needs NV*32 doubles.
template <int NV>
ZQ_D void block_sum(double (&v)[NV], double* sm) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
  __syncthreads();
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) sm[i * 32 + w] = v[i];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    double s = 0.0;
    for (int j = 0; j < nw; ++j) s += sm[i * 32 + j];
    v[i] = s;
  }
}
#endif

#define ZQ_CUDA_CHECK(expr)                                             \
  do {                                                                  \
    cudaError_t _e = (expr);                                            \
    if (_e != cudaSuccess) return zq_cuda_fail(_e, __FILE__, __LINE__); \
  } while (0)

int zq_cuda_fail(cudaError_t e, const char* file, int line);
