// Host-callable launchers of the zquatev B200 kernels (all asynchronous on `st`).
// Kernel ids K1..K10 follow SURVEY.md 7.2 / DESIGN.md.
#pragma once
#include <atomic>
#include "common.cuh"

namespace zq {

// ---- geometry of the K1 tiles (shared by matvec.cu and panel.cu) ----
constexpr int MV_TR = 128;   // tile rows  (4 rows per lane)
constexpr int MV_TC = 64;    // tile cols  (8 warps x 8 columns)
// panel inner products (W^H v, V^H v) fused into the K1 launch: rows [s, n) are cut into at most DOT_MAX_CHUNKS
// chunks of dot_chunk_rows(n - s) rows and the panel columns into groups of 8 (one warp per column), one CTA per
// (chunk, group) -- a short column still spreads over many SMs (a fixed 512-row chunk left ONE CTA with the whole
// panel at m = 512: K1 took 13 us at panel column 0 and 49 us at column 60).
constexpr int DOT_MIN_ROWS = 128, DOT_MAX_CHUNKS = 32;
__host__ __device__ inline int dot_chunk_rows(int rows) {
  int r = DOT_MIN_ROWS;
  while ((rows + r - 1) / r > DOT_MAX_CHUNKS) r *= 2;
  return r;
}
constexpr int K1_TPB_MAX = 8;      // column blocks per K1 CTA (a run); k1_tpb(m) picks 1, 2 or 4 from the trailing size
int k1_tpb(int m, int world);
constexpr int ROWS_PER_CTA = 256;
constexpr int PANEL_ROWS = 32;      // rows per CTA of the latency-bound panel kernels
constexpr int MAX_NB_PANEL = 64;    // largest panel width

// cudaFuncSetAttribute applies to the CURRENT device only, so the "already done" flags are per device
inline bool first_use_on_this_device(std::atomic<unsigned long long>& mask) {
  int dev = 0;
  cudaGetDevice(&dev);
  const unsigned long long bit = 1ull << (dev & 63);
  return !(mask.fetch_or(bit) & bit);
}

// launch of a kernel of the per-column chain: with ZQ_PDL != 0 (default) the launch carries the programmatic-stream-
// serialization attribute, see pdl_enter() in common.cuh
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_chain_smem(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_chain(void (*kern)(KArgs...), dim3 grid, dim3 block, cudaStream_t st, Args&&... args) {
  return launch_chain_smem(kern, grid, block, 0, st, static_cast<Args&&>(args)...);
}

struct PanelWs {
  int n, nb;
  size_t lda;         // leading dimension of A (complex elements)
  cplx* A;            // 2n x n: D rows [0,n), E rows [n,2n)
  cplx* pan;          // [4][nb][n]: Va, Vb, Wa, Wb
  quat* x;            // [xrec + 3] current (updated) column k in rows [k+1, n), followed at x[xrec] by the scalar
                      //   record of the column: (d_k, e_k = ||x||, tau_k, 0), alpha_k, u1^{-1}.  The reflector is
                      //   v[r] = x[r] u1^{-1} (v[k+1] = 1): every consumer forms it on the fly.
  int xrec;           // index of the record (n on one GPU; the capacity of the landing buffer with the peer exchange)
  unsigned int* counter;   // CTA arrival counter of col_update's last-block epilogue (always left at 0)
  const cplx* apanel; // multi-GPU: landing buffer of the current panel's columns [2 parts][MAX_NB_PANEL][apanel_ld]
  size_t apanel_ld;   //   (null on one GPU: col_update reads column k from A)
  quat* p;            // [n] tau * (M v - corrections)
  quat* pd;           // [ceil(n/MV_TC)][n]  direct partial sums of K1
  quat* pt;           // [ceil(n/MV_TR)][n]  transposed partial sums of K1
  quat* dotW;         // [nchunks][nb]  partial W^H v
  quat* dotV;         // [nchunks][nb]  partial V^H v
  double* nrm_part;   // [ceil(n/256)]
  double* g_part;     // [ceil(n/256)]
  double* d;          // [n]   diagonal of the real tridiagonal
  double* e;          // [n]   |alpha_k|
  double* tau;        // [n]
  quat* alpha;        // [n]   quaternion sub-diagonal
  quat* G;            // [n][nb] saved V^H v_i (strict upper part of the panel Gram matrix)
  // multi-GPU (1-D block-cyclic by 64-column blocks): this rank owns column blocks
  // J with J % world == rank.  world == 1: single GPU.
  int rank, world;
};

// Peer-memory exchange for the multi-GPU reduction (one process per GPU, buffers opened through
// CUDA IPC, NVLink/NVSwitch stores): the two per-column collectives -- broadcast of the reflector,
// all-reduce of the partial mat-vec -- are done by the panel kernels themselves with direct peer
// stores and sequence-number flags instead of NCCL calls between kernels.
constexpr int PX_MAXW = 8;
struct PeerX {
  int rank, world;
  size_t nmax;                              // capacity (quaternions per vector)
  int rbmax;                                // row blocks (of PANEL_ROWS rows) per vector: flag capacity
  cplx* apanel[PX_MAXW];                    // [g]: rank g's landing buffer for the current panel's columns: [2 parts][64][nmax]
  quat* ypart[PX_MAXW];                     // [g]: rank g's staging [2 parities][world][nmax]
  unsigned long long* flags[PX_MAXW];       // [g]: rank g's flags: [0] panel seq; [8 + ((parity*rbmax + rowblock)*PX_MAXW + src)] partial-y seq
  int* info;                                // local status word (bit 8 = exchange timeout)
};

// K2/K3 panel column kernels (panel.cu)
void launch_col_update(const PanelWs& w, int k, int j0, cudaStream_t st);
void launch_reduce_correct(const PanelWs& w, int k, int j0, cudaStream_t st);
// NCCL transport: (1) y_local = sum of this rank's K1 partials -> w.p rows [k+1, n);
// (2) after the all-reduce of w.p: corrections, tau scaling, partial Re(v^H p), unpack of the column
void launch_reduce_partial(const PanelWs& w, int k, cudaStream_t st);
void launch_correct(const PanelWs& w, int k, int j0, cudaStream_t st);
// multi-GPU, once per panel: the owner pushes rows [j0, n) of the panel's kb columns of D and E into every rank's landing
// buffer and publishes seq (peer exchange); the others wait for it.  NCCL transport: pack into a local buffer instead.
void launch_push_panel(const PanelWs& w, const PeerX& px, int j0, int kb, unsigned long long seq, cudaStream_t st);
void launch_wait_panel(const PeerX& px, unsigned long long seq, cudaStream_t st);
void launch_pack_panel(const PanelWs& w, cplx* buf, size_t ld, int j0, int kb, cudaStream_t st);
// peer-exchange variant of reduce_correct (seq = monotonically increasing sequence number of this column):
//   partial M v pushed per row block into every rank's staging slot + per-row-block flags, ordered sum, correction
void launch_reduce_correct_px(const PanelWs& w, const PeerX& px, int k, int j0, unsigned long long seq, cudaStream_t st);
void launch_finish_w(const PanelWs& w, int k_last, int j0, cudaStream_t st);
// K1 quaternion-Hermitian mat-vec on the lower triangles (+ fused panel dot products)
void launch_matvec(const PanelWs& w, int k, int j0, cudaStream_t st);
// stand-alone K1 for tests/bench: y = M[s:,s:] v with v = w.x[s..n) as given (record u1^{-1} must be 1)
void launch_matvec_only(const PanelWs& w, int s, quat* y, cudaStream_t st, bool gather = true);

// K5 (small.cu): the whole reduction of one matrix with n <= small_n_max() in one launch of one CTA (same outputs
// as the K1-K4 chain: d, e, tau, alpha, reflector tails in A, Gram columns G).  small_n_max() = ZQ_SMALL_N (read at
// every solve; default and upper limit SMALL_N_MAX, 0 = always use the chain); small_prepare() opts the kernel into
// its shared-memory size (once per device, outside any stream capture).
constexpr int SMALL_N_MAX = 256;
int small_n_max();
cudaError_t small_prepare();
void launch_tridiag_small(const PanelWs& w, cudaStream_t st);

// K4 operands: L (2m x 4kb, ld 2m) and R (m x 4kb, ld m) from the panel, m = n - r0
void launch_build_LR(const PanelWs& w, int r0, int kb, cplx* L, cplx* R, cudaStream_t st);

// K4/K6 complex GEMM:  C = alpha * op(A) * op(B) + beta * C
//   ta/tb: 0 = as stored, 1 = conjugate transpose.  lower != 0: only tiles with row >= col are
//   computed and, inside diagonal tiles, only entries with row >= col are stored.
//   Batched over `batch` with element strides sA, sB, sC.
void launch_zgemm(int ta, int tb, int M, int N, int K, cplx alpha, const cplx* A, size_t lda, const cplx* B,
                  size_t ldb, cplx beta, cplx* C, size_t ldc, int lower, int batch, size_t sA, size_t sB,
                  size_t sC, cudaStream_t st);
// split-K form (alpha = 1, beta = 0): the K range is cut into `nseg` segments (operand offsets segA, segB elements
// apart; K = length of one segment) of `chunks` pieces of kc each; piece z = seg*chunks + c writes its own partial
// product to Cparts + z*sC.  The caller sums the parts in a fixed order (launch_sum_parts): deterministic, no atomics.
// Used for Y = Phi(V)^H X of the back-transformation, whose 128 x ncols output gives too few CTAs on its own.
struct SplitK { int chunks = 0, kc = 0; size_t segA = 0, segB = 0; };
void launch_zgemm_splitk(int ta, int tb, int M, int N, int K, const cplx* A, size_t lda, const cplx* B, size_t ldb,
                         cplx* Cparts, size_t ldc, size_t sC, int nseg, const SplitK& sk, cudaStream_t st);
// Y[i] = sum_z parts[z*stride + i], i < count (complex), z ascending
void launch_sum_parts(size_t count, int nparts, const cplx* parts, size_t stride, cplx* Y, cudaStream_t st);
// 3M complex product (three real DMMA products per complex product, as the reference's zgemm3m) on/off.
// The driver enables it for n >= 1024, where its 25 % saving matters and the solver's residual is about
// half the reference's; below that the conventional four-product kernel keeps the last digit.
void zgemm_allow_3m(int on);
// same, restricted to the 64-column blocks cb0, cb0+cbs, cb0+2*cbs, ... (ncb of them) of C
void launch_zgemm_cb(int ta, int tb, int M, int N, int K, cplx alpha, const cplx* A, size_t lda, const cplx* B,
                     size_t ldb, cplx beta, cplx* C, size_t ldc, int lower, int batch, size_t sA, size_t sB,
                     size_t sC, int cb0, int cbs, int ncb, cudaStream_t st);

// K4/K6 quaternion GEMM with eight real products per quaternion product (qgemm.cu):  C = beta C + alpha op(A) op(B),
// alpha and beta real.  A quaternion matrix is a pair of complex column-major arrays, the b-part `off` elements behind
// the a-part.  ta/tb: 1 = the operand is the quaternion conjugate transpose of the stored array (stored K x M / N x K).
// lower: only entries with row >= col.  Either `batch` independent products (strides sA, sB, sC) or split-K (sk != null,
// sk->chunks pieces of sk->kc along K, piece z written to C + z sC; segA/segB unused).
void launch_qgemm(int ta, int tb, int M, int N, int K, double alpha, const cplx* A, size_t lda, size_t aoff, const cplx* B,
                  size_t ldb, size_t boff, double beta, cplx* C, size_t ldc, size_t coff, int lower, int batch, size_t sA,
                  size_t sB, size_t sC, const SplitK* sk, cudaStream_t st, int cb0 = 0, int cbs = 1, int ncb = -1);
// Development variant (qgemm8x.cu, ZQ_Q8X=1) for products whose operands are both panels (K4, the update half of K6): the
// eight component sums of every operand are formed once per panel by an elementwise pass (A8, B8: workspaces of
// qgemm_x_operand_doubles(rows, K) doubles), the product itself has no DADD in its main loop.  ta = 0 only.
bool qgemm_x_enabled();
size_t qgemm_x_operand_doubles(int rows_max, int kmax);
void launch_qgemm_x(int tb, int M, int N, int K, double alpha, const cplx* A, size_t lda, size_t aoff, const cplx* B, size_t ldb,
                    size_t boff, double beta, cplx* C, size_t ldc, size_t coff, int lower, double* A8, double* B8, cudaStream_t st);
// K4 operands of the quaternion form: Aq = [V W], Sq = [W V] (rows r0..n-1; each 2m x 2kb complex with the b-part
// stacked below the a-part, ld 2m), so that  M[r0:, r0:] -= Aq Sq^H
void launch_build_VW(const PanelWs& w, int r0, int kb, cplx* Aq, cplx* Sq, cudaStream_t st);

// K6 helpers (backtransform.cu)
//   P = Phi(V) of panel [j0, j0+kb): (2m x 2kb), ld 2m, m = n-1-j0 ; rows [0,m) <-> a-part rows j0+1..
void launch_build_phi(const PanelWs& w, int j0, int kb, cplx* P, cudaStream_t st);
//   two-panel form (ZQ_BT_PAIR): Phi(V) of a panel zero-padded to the m_op rows (per half) of the earlier panel it is
//   merged with, and T12 = [[Ta, 0], [0, Tb]] (the off-diagonal block -Ta (Pa^H Pb) Tb is added by GEMMs)
void launch_build_phi_padded(const PanelWs& w, int j0, int kb, cplx* P, int m_op, int row_off, cudaStream_t st);
void launch_assemble_T12(const cplx* Ta, int ka2, const cplx* Tb, int kb2, cplx* T12, cudaStream_t st);
//   quaternion form of the two-panel step: quaternion columns (Va; Vb) of a panel (zero-padded to m_op rows per half, starting
//   row_off rows down), and the Phi form of the merged T factor [[Ta_q, C_q], [0, Tb_q]] with C_q given stacked (2ka x kb)
void launch_build_vq(const PanelWs& w, int j0, int kb, cplx* Vq, int m_op, int row_off, cudaStream_t st);
void launch_assemble_T12q(const cplx* Ta, int ka, const cplx* Tb, int kb, const cplx* Cq, cplx* T12, cudaStream_t st);
//   T of every panel (panel j at Tall + j * 4 nb^2: 2kb x 2kb complex, ld 2kb) from saved Gram columns G and tau
void launch_build_T_all(const PanelWs& w, cplx* Tall, cudaStream_t st);
//   phase chain s (n quats) from alpha; X0 = diag(s) Z[:, perm]
void launch_phase_chain(int n, const quat* alpha, const double* e, quat* s, cudaStream_t st);
void launch_scale_Z(int n, int ncols, const double* Z, size_t ldz, const int* perm, const quat* s, cplx* X, size_t ldx,
                    cudaStream_t st);
// K10 pairing, in place: X sits in the RIGHT half (columns n..2n-1); on exit the left
// half holds (U;V) = X and the right half Theta(X) = (-conj V; conj U)
void launch_swap_pairing(int n, int ncols, cplx* Out, size_t ld, cudaStream_t st);
// right half (columns c..2c-1) of a structured 2r x 2c array from its left half
void launch_fill_pairing(int r, int c, cplx* M, size_t ld, cudaStream_t st);
// R (2n x ncols block of X) <- Theta(R) in place (host-pointer pipeline: X itself has been downloaded already)
void launch_theta_inplace(int n, int ncols, cplx* R, size_t ld, cudaStream_t st);

// K8 tridiagonal divide & conquer (dc.cu).  d,e: n ; Z: n x n (ldz) ; on exit w ascending = d[perm[.]]
struct DcWs;
DcWs* dc_create(int n);
void dc_destroy(DcWs*);
size_t dc_bytes(int n);
long dc_launches(const DcWs* ws);
// returns 0 / cuda error; eigenvalues (ascending) -> wout[n]; eigenvectors: column j of the result is
// column perm[j] of Zres (pointer returned in *Zres, ld n).  info flag on device: *info_dev != 0 -> failure.
// multi-GPU: the top-level merge GEMM is split by eigenvector column blocks; `allgather(buf, count, rank, st)`
// gathers `count` doubles per rank in place (buf + rank*count is this rank's part)
struct DcDist { int rank, world; int (*allgather)(double* buf, size_t count, int rank, cudaStream_t st); };
int dc_solve(DcWs* ws, int n, const double* d, const double* e, double* wout, double** Zres, int** perm,
             int* info_dev, cudaStream_t st, const DcDist* dd = nullptr);

// K9 eigenvalues only: Sturm bisection
// scratch: n + 3 doubles
void launch_bisect(int n, const double* d, const double* e, double* w, double* scratch, cudaStream_t st, int jlo = 0, int jhi = -1);

// input scaling (scale.cu): largest entry brought into [sqrt(safmin/eps), sqrt(eps/safmin)] like LAPACK's zheev driver
size_t scale_scratch_doubles(int n);
void launch_scale_input(cplx* A, size_t lda, int n, double* scratch, cudaStream_t st);
void launch_unscale_eig(int n, double* eig, const double* scratch, cudaStream_t st);

// misc
// sets bit 0 of *flag when d / e hold NaN or Inf, and then zeroes them (the eigensolver must not see NaNs)
void launch_check_finite(int n, double* d, double* e, int* flag, cudaStream_t st);

}  // namespace zq
