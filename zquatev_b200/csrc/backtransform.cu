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

// Same as k_build_phi for a panel that is merged with an EARLIER panel (two-panel back-transformation, solver.cu):
// the operand has m_op rows per half (those of the earlier panel), this panel starts `row_off` rows further down,
// the rows above it are zero.
__global__ void __launch_bounds__(256) k_build_phi_padded(PanelWs w, int j0, int kb, cplx* P, int m_op, int row_off) {
  const int n = w.n;
  const int t = blockIdx.y;
  const int ro = blockIdx.x * 256 + threadIdx.x;      // row inside the operand's half
  if (ro >= m_op) return;
  const int rr = ro - row_off;                        // row inside this panel's own half (row j0+1+rr of the matrix)
  cplx va = cmake(0, 0), vb = cmake(0, 0);
  if (rr == t) va = cmake(1, 0);
  else if (rr > t) {
    const size_t r = (size_t)(j0 + 1 + rr), k = (size_t)(j0 + t);
    va = w.A[r + k * w.lda];
    vb = w.A[n + r + k * w.lda];
  }
  const size_t ld = 2 * (size_t)m_op;
  P[ro + (size_t)t * ld] = va;
  P[m_op + ro + (size_t)t * ld] = vb;
  P[ro + (size_t)(kb + t) * ld] = cneg(cconj(vb));
  P[m_op + ro + (size_t)(kb + t) * ld] = cconj(va);
}

// Quaternion columns only (no Theta image): Vq[:, t] = (Va_t; Vb_t) of panel j0, a-part rows [0, m_op), b-part rows
// [m_op, 2 m_op), this panel starting `row_off` rows down (zeros above).  Two calls into adjacent column ranges build the
// merged panel [V_ja | V_jb] of the two-panel back-transformation on the quaternion GEMM.
__global__ void __launch_bounds__(256) k_build_vq(PanelWs w, int j0, cplx* Vq, int m_op, int row_off) {
  const int n = w.n;
  const int t = blockIdx.y;
  const int ro = blockIdx.x * 256 + threadIdx.x;
  if (ro >= m_op) return;
  const int rr = ro - row_off;
  cplx va = cmake(0, 0), vb = cmake(0, 0);
  if (rr == t) va = cmake(1, 0);
  else if (rr > t) {
    const size_t r = (size_t)(j0 + 1 + rr), k = (size_t)(j0 + t);
    va = w.A[r + k * w.lda];
    vb = w.A[n + r + k * w.lda];
  }
  const size_t ld = 2 * (size_t)m_op;
  Vq[ro + (size_t)t * ld] = va;
  Vq[m_op + ro + (size_t)t * ld] = vb;
}

// Phi form (2 Kq x 2 Kq complex, ld 2 Kq, rows / columns ordered [a-parts (Kq) | b-parts (Kq)]) of the merged quaternion T factor
//   T12_q = [[Ta_q, C_q], [0, Tb_q]],   Kq = ka + kb,
// from the per-panel Phi forms Ta (2ka x 2ka), Tb (2kb x 2kb) -- whose first columns hold (T_a; T_b) stacked -- and the cross
// block C_q = -Ta_q (Va^H Vb) Tb_q given stacked (2ka x kb, ld 2ka).
__global__ void __launch_bounds__(256) k_assemble_T12q(const cplx* __restrict__ Ta, int ka, const cplx* __restrict__ Tb, int kb,
                                                       const cplx* __restrict__ Cq, cplx* __restrict__ T12) {
  const int Kq = ka + kb, ld = 2 * Kq;
  for (int idx = blockIdx.x * 256 + threadIdx.x; idx < Kq * Kq; idx += gridDim.x * 256) {
    const int r = idx % Kq, c = idx / Kq;
    cplx a = cmake(0, 0), b = cmake(0, 0);
    if (r < ka && c < ka) { a = Ta[r + (size_t)c * 2 * ka]; b = Ta[ka + r + (size_t)c * 2 * ka]; }
    else if (r >= ka && c >= ka) { a = Tb[(r - ka) + (size_t)(c - ka) * 2 * kb]; b = Tb[kb + (r - ka) + (size_t)(c - ka) * 2 * kb]; }
    else if (r < ka && c >= ka) { a = Cq[r + (size_t)(c - ka) * 2 * ka]; b = Cq[ka + r + (size_t)(c - ka) * 2 * ka]; }
    T12[r + (size_t)c * ld] = a;
    T12[Kq + r + (size_t)c * ld] = b;
    T12[r + (size_t)(Kq + c) * ld] = cneg(cconj(b));
    T12[Kq + r + (size_t)(Kq + c) * ld] = cconj(a);
  }
}

// T12 (Kc x Kc, Kc = 2ka + 2kb, ld Kc) <- [[Ta, 0], [0, Tb]]; the upper-right block is filled by a GEMM afterwards
__global__ void __launch_bounds__(256) k_assemble_T12(const cplx* __restrict__ Ta, int ka2, const cplx* __restrict__ Tb, int kb2,
                                                      cplx* __restrict__ T12) {
  const int Kc = ka2 + kb2;
  for (int idx = blockIdx.x * 256 + threadIdx.x; idx < Kc * Kc; idx += gridDim.x * 256) {
    const int r = idx % Kc, c = idx / Kc;
    cplx v = cmake(0, 0);
    if (r < ka2 && c < ka2) v = Ta[r + (size_t)c * ka2];
    else if (r >= ka2 && c >= ka2) v = Tb[(r - ka2) + (size_t)(c - ka2) * kb2];
    T12[idx] = v;
  }
}

// T = Phi(T_q), T_q upper triangular quaternion kb x kb:
//   T_q[i,i] = tau_i ;  T_q[0:i, i] = -tau_i T_q[0:i,0:i] g_i ,  g_i = V[:, 0:i]^H v_i  (saved in G)
// so that H_{j0} ... H_{j0+kb-1} = I - V T_q V^H (forward, column-wise; zlarft analogue).
// One CTA per panel (blockIdx.x = panel index, its T at T + blockIdx.x * 4 nb^2): the column recurrence is serial
// inside a panel (64 dependent steps, ~180 us), but G and tau of ALL panels are known once the reduction is done,
// so every T is built by one launch before the back-transformation instead of one launch per panel on its
// critical path (11 ms of 80 ms at n = 4096).
__global__ void __launch_bounds__(128) k_build_T(PanelWs w, cplx* Tall) {
  const int j0 = blockIdx.x * w.nb;
  const int kb = min(w.nb, w.n - 1 - j0);
  if (kb <= 0) return;
  cplx* T = Tall + (size_t)blockIdx.x * 4 * w.nb * w.nb;
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
// Ordered prefix product of unit quaternions (associative, not commutative) as a one-CTA scan: thread t owns a
// contiguous segment of len links, forms the segment product, the segment products are scanned across the CTA
// (Hillis-Steele in shared memory, later factor on the LEFT), then every thread replays its segment on top of
// the product of all earlier segments.  Each s_k is renormalised, so rounding cannot drift off the unit sphere.
constexpr int PC_NT = 1024;
ZQ_D quat pc_link(const quat* alpha, const double* e, int k) {
  const double ek = e[k];
  return ek > 0.0 ? qscale(alpha[k], 1.0 / ek) : qmake(cmake(1, 0), cmake(0, 0));
}
ZQ_D quat pc_unit(quat q) { return qscale(q, rsqrt(qnorm2(q))); }
__global__ void __launch_bounds__(PC_NT) k_phase_chain(int n, const quat* __restrict__ alpha, const double* __restrict__ e,
                                                       quat* __restrict__ s) {
  __shared__ quat sc[PC_NT];
  const int t = threadIdx.x, links = n - 1;
  const int len = (links + PC_NT - 1) / PC_NT;
  const int k0 = t * len, k1 = min(links, k0 + len);
  quat seg = qmake(cmake(1, 0), cmake(0, 0));
  for (int k = k0; k < k1; ++k) seg = pc_unit(qmul(pc_link(alpha, e, k), seg));
  sc[t] = seg;
  __syncthreads();
  for (int off = 1; off < PC_NT; off <<= 1) {
    quat mine = sc[t];
    if (t >= off) mine = pc_unit(qmul(mine, sc[t - off]));
    __syncthreads();
    sc[t] = mine;
    __syncthreads();
  }
  quat cur = t > 0 ? sc[t - 1] : qmake(cmake(1, 0), cmake(0, 0));   // product of all earlier segments = s[k0]
  if (t == 0) s[0] = cur;
  for (int k = k0; k < k1; ++k) {
    cur = pc_unit(qmul(pc_link(alpha, e, k), cur));
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
  int src = perm[j];
  if ((unsigned)src >= (unsigned)n) src = j;          // a failed eigensolver (info > 0) must not turn into a wild read
  const double z = Z[(size_t)r + (size_t)src * ldz];
  const quat q = s[r];
  X[(size_t)r + (size_t)j * ldx] = cscale(q.a, z);
  X[(size_t)(n + r) + (size_t)j * ldx] = cscale(q.b, z);
}

// K10: in-place pairing: X sits in the right half; left half <- X, right half <- Theta(X) = (-conj(V); conj(U))   (zquatev.cc:93-98)
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

// K10, host-pointer pipeline: X (in the right half) has already been downloaded into the caller's LEFT half; turn it into
// Theta(X) in place for the download into the caller's right half.  The left half of the device array (the reflector
// storage the remaining column blocks still need) is not touched.
__global__ void __launch_bounds__(256) k_theta_inplace(int n, cplx* R, size_t ld) {
  const int r = blockIdx.x * 256 + threadIdx.x;
  if (r >= n) return;
  cplx* c = R + (size_t)blockIdx.y * ld;
  const cplx u = c[r], v = c[n + r];
  c[r] = cneg(cconj(v));
  c[n + r] = cconj(u);
}

// right half from the left half of a structured 2r x 2c array (caller-side helper of zquatev_b200_fill_pairing)
__global__ void __launch_bounds__(256) k_fill_pairing(int r_, int c_, cplx* M, size_t ld) {
  const int r = blockIdx.x * 256 + threadIdx.x;
  if (r >= r_) return;
  const cplx* L = M + (size_t)blockIdx.y * ld;
  cplx* R = M + (size_t)(c_ + blockIdx.y) * ld;
  const cplx u = L[r], v = L[r_ + r];
  R[r] = cneg(cconj(v));
  R[r_ + r] = cconj(u);
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
// A non-finite tridiagonal (NaN / Inf in the input) is reported through info; the eigensolver behind it then runs on
// a zero matrix instead of NaNs, whose comparisons would leave its index arrays (ranks, permutations) undefined.
__global__ void k_sanitize(int n, double* d, double* e, const int* flag) {
  if (!(*flag & 1)) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { d[i] = 0.0; e[i] = 0.0; }
}

}  // namespace

void launch_build_phi(const PanelWs& w, int j0, int kb, cplx* P, cudaStream_t st) {
  const int m = w.n - 1 - j0;
  dim3 g((m + 255) / 256, kb);
  k_build_phi<<<g, 256, 0, st>>>(w, j0, kb, P);
}

void launch_build_phi_padded(const PanelWs& w, int j0, int kb, cplx* P, int m_op, int row_off, cudaStream_t st) {
  dim3 g((m_op + 255) / 256, kb);
  k_build_phi_padded<<<g, 256, 0, st>>>(w, j0, kb, P, m_op, row_off);
}

void launch_build_vq(const PanelWs& w, int j0, int kb, cplx* Vq, int m_op, int row_off, cudaStream_t st) {
  dim3 g((m_op + 255) / 256, kb);
  k_build_vq<<<g, 256, 0, st>>>(w, j0, Vq, m_op, row_off);
}

void launch_assemble_T12q(const cplx* Ta, int ka, const cplx* Tb, int kb, const cplx* Cq, cplx* T12, cudaStream_t st) {
  const int Kq = ka + kb;
  k_assemble_T12q<<<(Kq * Kq + 255) / 256, 256, 0, st>>>(Ta, ka, Tb, kb, Cq, T12);
}

void launch_assemble_T12(const cplx* Ta, int ka2, const cplx* Tb, int kb2, cplx* T12, cudaStream_t st) {
  const int Kc = ka2 + kb2;
  k_assemble_T12<<<(Kc * Kc + 255) / 256, 256, 0, st>>>(Ta, ka2, Tb, kb2, T12);
}

void launch_build_T_all(const PanelWs& w, cplx* Tall, cudaStream_t st) {
  const int npanels = (w.n - 1 + w.nb - 1) / w.nb;
  if (npanels > 0) k_build_T<<<npanels, 128, 0, st>>>(w, Tall);
}

void launch_phase_chain(int n, const quat* alpha, const double* e, quat* s, cudaStream_t st) {
  k_phase_chain<<<1, PC_NT, 0, st>>>(n, alpha, e, s);
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

void launch_swap_pairing(int n, int ncols, cplx* Out, size_t ld, cudaStream_t st) {
  if (ncols <= 0) return;
  dim3 g((n + 255) / 256, ncols);
  k_swap_pairing<<<g, 256, 0, st>>>(n, Out, ld);
}

void launch_fill_pairing(int r, int c, cplx* M, size_t ld, cudaStream_t st) {
  dim3 g((r + 255) / 256, c);
  k_fill_pairing<<<g, 256, 0, st>>>(r, c, M, ld);
}

void launch_theta_inplace(int n, int ncols, cplx* R, size_t ld, cudaStream_t st) {
  if (ncols <= 0) return;
  dim3 g((n + 255) / 256, ncols);
  k_theta_inplace<<<g, 256, 0, st>>>(n, R, ld);
}

void launch_check_finite(int n, double* d, double* e, int* flag, cudaStream_t st) {
  k_check_finite<<<(n + 255) / 256, 256, 0, st>>>(n, d, e, flag);
  k_sanitize<<<(n + 255) / 256, 256, 0, st>>>(n, d, e, flag);
}

}  // namespace zq
