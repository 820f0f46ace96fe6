// K6 helpers, K7 layout kernels and K10 pairing.
//
// The reference accumulates Q = Phi(Q0,Q1) forward during the reduction (blocked.cc:477-544,
// 48 n^3 flop) and then multiplies U = Q0 Z, V = Q1 Z (zquatev.cc:87-90, 16 n^3).  Here the
// reflectors stay in compact-WY form and are applied BACKWARD to X0 = diag(s) Z (32 n^3):
// per panel  Y = Phi(V)^H X,  X -= Phi(V) (T Y)  as plain complex GEMMs (zgemm.cu), which keeps
// the quaternion block structure because Phi is a *-homomorphism.  The layout shuffles of
// transpose.cc / supermat.h become the operand-staging kernels below.
#include "kernels.h"

namespace zq {
namespace {

// P = Phi(V) = [[Va, -conj(Vb)], [Vb, conj(Va)]]   (2m x 2kb, ld 2m), unit diagonal, zeros above.
__global__ void __launch_bounds__(256) k_build_phi(PanelWs w, int j0, int kb, cplx* P) {
  const int n = w.n, m = n - 1 - j0;
  const int t = blockIdx.y;
  const int rr = blockIdx.x * 256 + threadIdx.x;
  if (rr >= m) return;
  cplx va = cmake(0, 0), vb = cmake(0, 0);
  if (rr == t) va = cmake(1, 0);
  else if (rr > t) {
    const size_t r = (size_t)(j0 + 1 + rr), k = (size_t)(j0 + t);
    va = w.A[r + k * w.lda];
    vb = w.A[n + r + k * w.lda];
  }
  const size_t ld = 2 * (size_t)m;
  P[rr + (size_t)t * ld] = va;
  P[m + rr + (size_t)t * ld] = vb;
  P[rr + (size_t)(kb + t) * ld] = cneg(cconj(vb));
  P[m + rr + (size_t)(kb + t) * ld] = cconj(va);
}

// T = Phi(T_q), T_q upper triangular quaternion kb x kb:
//   T_q[i,i] = tau_i ;  T_q[0:i, i] = -tau_i T_q[0:i,0:i] g_i ,  g_i = V[:, 0:i]^H v_i  (saved in G)
// so that H_{j0} ... H_{j0+kb-1} = I - V T_q V^H (forward, column-wise; zlarft analogue).
__global__ void __launch_bounds__(128) k_build_T(PanelWs w, int j0, int kb, cplx* T) {
  const int ld = 2 * kb, t = threadIdx.x;
  for (int idx = t; idx < ld * ld; idx += blockDim.x) T[idx] = cmake(0, 0);
  __syncthreads();
  for (int i = 0; i < kb; ++i) {
    const double tau = w.tau[j0 + i];
    if (t < i) {
      quat acc = qzero();
      for (int u = t; u < i; ++u) {
        quat tq = qmake(T[t + (size_t)u * ld], T[kb + t + (size_t)u * ld]);
        qfma(acc, tq, w.G[(size_t)(j0 + i) * w.nb + u]);
      }
      acc = qscale(acc, -tau);
      T[t + (size_t)i * ld] = acc.a;
      T[kb + t + (size_t)i * ld] = acc.b;
      T[t + (size_t)(kb + i) * ld] = cneg(cconj(acc.b));
      T[kb + t + (size_t)(kb + i) * ld] = cconj(acc.a);
    } else if (t == i) {
      T[i + (size_t)i * ld] = cmake(tau, 0);
      T[kb + i + (size_t)(kb + i) * ld] = cmake(tau, 0);
    }
    __syncthreads();
  }
}

// s_0 = 1, s_{k+1} = (alpha_k/|alpha_k|) s_k : diag(s)^H T_q diag(s) is real symmetric.
__global__ void k_phase_chain(int n, const quat* alpha, const double* e, quat* s) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  quat cur = qmake(cmake(1, 0), cmake(0, 0));
  s[0] = cur;
  for (int k = 0; k + 1 < n; ++k) {
    const double ek = e[k];
    if (ek > 0.0) {
      cur = qmul(qscale(alpha[k], 1.0 / ek), cur);
      cur = qscale(cur, rsqrt(qnorm2(cur)));
    }
    s[k + 1] = cur;
  }
}

// X0 = diag(s) Z[:, perm]  (Z real n x n) into the stacked output [Xa; Xb]
__global__ void __launch_bounds__(256) k_scale_Z(int n, const double* __restrict__ Z, size_t ldz,
                                                 const int* __restrict__ perm, const quat* __restrict__ s,
                                                 cplx* __restrict__ X, size_t ldx) {
  const int j = blockIdx.y;
  const int r = blockIdx.x * 256 + threadIdx.x;
  if (r >= n) return;
  const double z = Z[(size_t)r + (size_t)perm[j] * ldz];
  const quat q = s[r];
  X[(size_t)r + (size_t)j * ldx] = cscale(q.a, z);
  X[(size_t)(n + r) + (size_t)j * ldx] = cscale(q.b, z);
}

// K10: right half = Theta(left half): columns n+j = (-conj(V_j); conj(U_j))   (zquatev.cc:93-98)
__global__ void __launch_bounds__(256) k_pairing(int n, cplx* Out, size_t ld) {
  const int j = blockIdx.y;
  const int r = blockIdx.x * 256 + threadIdx.x;
  if (r >= n) return;
  const cplx u = Out[(size_t)r + (size_t)j * ld];
  const cplx v = Out[(size_t)(n + r) + (size_t)j * ld];
  Out[(size_t)r + (size_t)(n + j) * ld] = cneg(cconj(v));
  Out[(size_t)(n + r) + (size_t)(n + j) * ld] = cconj(u);
}

__global__ void __launch_bounds__(256) k_swap_pairing(int n, cplx* Out, size_t ld) {
  const int j = blockIdx.y;
  const int r = blockIdx.x * 256 + threadIdx.x;
  if (r >= n) return;
  cplx* L = Out + (size_t)j * ld;
  cplx* R = Out + (size_t)(n + j) * ld;
  const cplx u = R[r], v = R[n + r];
  L[r] = u;
  L[n + r] = v;
  R[r] = cneg(cconj(v));
  R[n + r] = cconj(u);
}

// deterministic split-K reduction: Y = parts[0] + parts[1] + ... (fixed order)
__global__ void __launch_bounds__(256) k_sum_parts(size_t count, int nparts, const cplx* __restrict__ parts, size_t stride,
                                                   cplx* __restrict__ Y) {
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < count; i += (size_t)gridDim.x * 256) {
    cplx acc = parts[i];
    for (int z = 1; z < nparts; ++z) {
      const cplx t = parts[(size_t)z * stride + i];
      acc.x += t.x;
      acc.y += t.y;
    }
    Y[i] = acc;
  }
}

__global__ void k_check_finite(int n, const double* d, const double* e, int* flag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (!isfinite(d[i]) || (i + 1 < n && !isfinite(e[i]))) atomicOr(flag, 1);
}

}  // namespace

void launch_build_phi(const PanelWs& w, int j0, int kb, cplx* P, cudaStream_t st) {
  const int m = w.n - 1 - j0;
  dim3 g((m + 255) / 256, kb);
  k_build_phi<<<g, 256, 0, st>>>(w, j0, kb, P);
}

void launch_build_T(const PanelWs& w, int j0, int kb, cplx* T, cudaStream_t st) {
  k_build_T<<<1, 128, 0, st>>>(w, j0, kb, T);
}

void launch_phase_chain(int n, const quat* alpha, const double* e, quat* s, cudaStream_t st) {
  k_phase_chain<<<1, 32, 0, st>>>(n, alpha, e, s);
}

void launch_scale_Z(int n, int ncols, const double* Z, size_t ldz, const int* perm, const quat* s, cplx* X, size_t ldx,
                    cudaStream_t st) {
  if (ncols <= 0) return;
  dim3 g((n + 255) / 256, ncols);
  k_scale_Z<<<g, 256, 0, st>>>(n, Z, ldz, perm, s, X, ldx);
}

void launch_sum_parts(size_t count, int nparts, const cplx* parts, size_t stride, cplx* Y, cudaStream_t st) {
  if (count == 0 || nparts <= 0) return;
  size_t blocks = (count + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  k_sum_parts<<<(unsigned)blocks, 256, 0, st>>>(count, nparts, parts, stride, Y);
}

void launch_pairing(int n, cplx* Out, size_t ld, cudaStream_t st) {
  dim3 g((n + 255) / 256, n);
  k_pairing<<<g, 256, 0, st>>>(n, Out, ld);
}

void launch_swap_pairing(int n, int ncols, cplx* Out, size_t ld, cudaStream_t st) {
  if (ncols <= 0) return;
  dim3 g((n + 255) / 256, ncols);
  k_swap_pairing<<<g, 256, 0, st>>>(n, Out, ld);
}

void launch_check_finite(int n, const double* d, const double* e, int* flag, cudaStream_t st) {
  k_check_finite<<<(n + 255) / 256, 256, 0, st>>>(n, d, e, flag);
}

}  // namespace zq
