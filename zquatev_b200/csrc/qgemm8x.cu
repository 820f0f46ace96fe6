// Quaternion GEMM with eight real products on PRE-COMBINED operands (ZQ_Q8X=1; development variant of qgemm.cu for the two
// products whose operands are both skinny panels: the trailing update K4, (D + jE) -= [V W] [W V]^H, and the update half of the
// back-transformation K6, X -= V (T Y)).
//
// k_qgemm8 (qgemm.cu) forms the eight +/- sums of the four component planes of every operand fragment in registers, inside
// the main loop: one DADD per DMMA on the SAME physical FP64 pipe (ceiling 0.86 of the DMMA rate,
// profiles/r02_dmma_dadd_coissue.jsonl), 64 operand registers next to the 128 accumulator registers (~200 registers: two
// CTAs = 8 warps per SM, issue stalls dominated by fixed-latency waits, profiles/r02_ncu_qgemm.md).  Where BOTH operands are
// panels that every output tile re-reads (K <= 128 quaternions), the sums can be formed ONCE per panel by an elementwise
// pass: operand e of product e is then a plain real matrix
//     A8[e][k][m]  (m contiguous, rows padded with zeros to a multiple of 32, k padded to a multiple of 8)
//     B8[e][k][n]  (n contiguous, same padding)
// and the GEMM is eight real DMMA products with eight accumulator sets and the usual recombination in the epilogue:
// no DADD in the main loop, 8-byte fragment loads straight from shared memory (conflict-free: plane row stride 36 doubles),
// no bounds predicate anywhere in the loop (the padding is zeros), ~40 registers less, three CTAs (12 warps) per SM.
// The price is twice the operand bytes from L2 to shared memory (ncu: lts throughput of k_qgemm8 is 11-17 % of peak).
//
// Tiling: CTA = 4 warps (2 x 2), warp tile 16 x 16 quaternions, CTA tile 32 x 32, BK = 8, two cp.async stages of 36 KB,
// persistent CTAs (3 per SM) whose operand pipeline runs across tile boundaries, C read-modify-written in the epilogue
// (its lines are prefetched into L2 when the tile starts).
#include "kernels.h"
#include "gemm_tiles.cuh"

namespace zq {
namespace {

constexpr int XBM = 32, XBN = 32, XBK = 8, XNT = 128, XST = 2;
constexpr int XLD = 36;                    // doubles per k-row of a plane tile (32 + 4: the 16 lanes of a half warp, 4 rows x 4
                                           // consecutive doubles, hit 16 distinct 8-byte banks)
constexpr int XPLANE = XBK * XLD;          // doubles per plane per stage
constexpr int XOP = 8 * XPLANE;            // doubles per operand per stage (18 KB)
constexpr int XSTAGE = 2 * XOP;            // doubles per stage (36 KB)

struct XArgs {
  int M, N, K8;                            // K8: padded K (multiple of XBK)
  double alpha, beta;
  const double* A8; int ldA;               // A8[(e K8 + k) ldA + m]
  const double* B8; int ldB;               // B8[(e K8 + k) ldB + n]
  cplx* C; size_t ldc, coff;
  int lower;
};

// ---- operand pre-combination --------------------------------------------------------------------------------------
// A(m, k) = (SA[m + k lda], SA[aoff + m + k lda]) as stored (TA = 0 only)
__global__ void __launch_bounds__(256) k_combine_a(const cplx* __restrict__ A, size_t lda, size_t aoff, int M, int K, double* __restrict__ A8,
                                                   int ldA, int K8) {
  const int m = blockIdx.x * 256 + threadIdx.x, k = blockIdx.y;
  if (m >= ldA) return;
  double al[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (m < M && k < K) combos_a<false>(A[(size_t)m + (size_t)k * lda], A[aoff + (size_t)m + (size_t)k * lda], al);
#pragma unroll
  for (int e = 0; e < 8; ++e) A8[((size_t)e * K8 + k) * ldA + m] = al[e];
}
// TB = 0: B(k, n) = SB[k + n ldb];  TB = 1: B(k, n) = conj_q(SB[n + k ldb])
template <int TB>
__global__ void __launch_bounds__(256) k_combine_b(const cplx* __restrict__ B, size_t ldb, size_t boff, int N, int K, double* __restrict__ B8,
                                                   int ldB, int K8) {
  const int n = blockIdx.x * 256 + threadIdx.x, k = blockIdx.y;
  if (n >= ldB) return;
  double be[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (n < N && k < K) {
    const size_t at = TB ? (size_t)n + (size_t)k * ldb : (size_t)k + (size_t)n * ldb;
    combos_b<TB == 1>(B[at], B[boff + at], be);
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) B8[((size_t)e * K8 + k) * ldB + n] = be[e];
}

// ---- the product ----------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(XNT, 3) k_qgemm8x(XArgs p, int tiles_m, int tiles_n) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* smem = reinterpret_cast<double*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = (warp & 1) * 16, wn = (warp >> 1) * 16;
  const int g = lane >> 2, q = lane & 3;
  const bool bzero = (p.beta == 0.0);
  const int nk = p.K8 / XBK;
  const int T = tiles_m * tiles_n;
  auto tile_rc = [&](int t, int& r0, int& c0) -> bool {
    r0 = (t % tiles_m) * XBM;
    c0 = (t / tiles_m) * XBN;
    return !(p.lower && r0 + XBM - 1 < c0);
  };
  auto next_valid = [&](int t) -> int {
    int r0, c0;
    while (t < T && !tile_rc(t, r0, c0)) t += gridDim.x;
    return t;
  };

  // loader: thread = (16-byte chunk m2 along the tile dimension, k row kk of the stage); one copy per plane and operand
  const int m2 = tid & 15, kk = tid >> 4;
  const size_t planeA = (size_t)p.K8 * p.ldA, planeB = (size_t)p.K8 * p.ldB;
  const int sdst = kk * XLD + 2 * m2;
  int lt = next_valid((int)blockIdx.x), lkt = 0;
  unsigned gl = 0;                                   // stages issued so far (slot = gl % XST)
  const double* gA = p.A8;
  const double* gB = p.B8;
  auto loader_init = [&]() {
    int r0, c0;
    tile_rc(lt, r0, c0);
    gA = p.A8 + (size_t)kk * p.ldA + r0 + 2 * m2;
    gB = p.B8 + (size_t)kk * p.ldB + c0 + 2 * m2;
  };
  auto issue_next = [&]() {
    if (lt < T) {
      double* s = smem + (size_t)(gl % XST) * XSTAGE + sdst;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        cp_async16(s + e * XPLANE, gA + (size_t)e * planeA, true);
        cp_async16(s + XOP + e * XPLANE, gB + (size_t)e * planeB, true);
      }
      gA += (size_t)XBK * p.ldA;
      gB += (size_t)XBK * p.ldB;
      if (++lkt == nk) {
        lkt = 0;
        lt = next_valid(lt + (int)gridDim.x);
        if (lt < T) loader_init();
      }
    }
    ++gl;
    cp_async_commit();
  };
  auto prefetch_c = [&](int t) {                     // warm L2 with the C tile the epilogue will read-modify-write
    int r0, c0;
    tile_rc(t, r0, c0);
    for (int e = tid; e < XBN * (XBM / 8) * 2; e += XNT) {
      const int half = e / (XBN * (XBM / 8)), f = e % (XBN * (XBM / 8));
      const int c = c0 + f / (XBM / 8), r = r0 + (f % (XBM / 8)) * 8;
      if (c < p.N && r < p.M) asm volatile("prefetch.global.L2 [%0];\n" ::"l"(p.C + (size_t)half * p.coff + (size_t)r + (size_t)c * p.ldc));
    }
  };

  int ct = lt;                                       // consumer cursor
  if (ct >= T) return;
  loader_init();
  if (!bzero) prefetch_c(ct);
#pragma unroll
  for (int s = 0; s < XST - 1; ++s) issue_next();
  unsigned gu = 0;                                   // stages consumed so far

  while (ct < T) {
    int r0, c0;
    tile_rc(ct, r0, c0);
    double acc[8][2][2][2];
#pragma unroll
    for (int e = 0; e < 8; ++e)
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) acc[e][i][j][0] = acc[e][i][j][1] = 0.0;
    for (int kt = 0; kt < nk; ++kt, ++gu) {
      cp_async_wait<XST - 2>();
      __syncthreads();                               // stage gu has landed; every thread is done with stage gu - 1
      issue_next();                                  // refills the slot of stage gu - 1
      const double* sA = smem + (size_t)(gu % XST) * XSTAGE;
      const double* sB = sA + XOP;
#pragma unroll
      for (int k4 = 0; k4 < XBK; k4 += 4) {
        const double* pa = sA + (k4 + q) * XLD + wm + g;
        const double* pb = sB + (k4 + q) * XLD + wn + g;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const double a0 = pa[e * XPLANE], a1 = pa[e * XPLANE + 8];
          const double b0 = pb[e * XPLANE], b1 = pb[e * XPLANE + 8];
          dmma(acc[e][0][0][0], acc[e][0][0][1], a0, b0);
          dmma(acc[e][0][1][0], acc[e][0][1][1], a0, b1);
          dmma(acc[e][1][0][0], acc[e][1][0][1], a1, b0);
          dmma(acc[e][1][1][0], acc[e][1][1][1], a1, b1);
        }
      }
    }
    // epilogue: recombine the eight products; lane holds rows wm+8i+g, columns wn+8j+2q+{0,1}
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c = c0 + wn + 8 * j + 2 * q + h;
        if (c >= p.N) continue;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int r = r0 + wm + 8 * i + g;
          if (r >= p.M || (p.lower && r < c)) continue;
          const double p1 = acc[0][i][j][h], p2 = acc[1][i][j][h], p3 = acc[2][i][j][h], p4 = acc[3][i][j][h];
          const double s123 = (p1 + p2) + p3;
          const double sh = 0.5 * (s123 + p4);
          const double q0 = (sh - p1) + acc[4][i][j][h];
          const double q1 = (sh - s123) + acc[5][i][j][h];
          const double q2 = (sh - p2) + acc[6][i][j][h];
          const double q3 = (sh - p3) + acc[7][i][j][h];
          cplx va = cmake(p.alpha * q0, p.alpha * q1);
          cplx vb = cmake(p.alpha * q2, -p.alpha * q3);
          cplx* ca = p.C + (size_t)r + (size_t)c * p.ldc;
          cplx* cb = ca + p.coff;
          if (!bzero) {
            const cplx oa = *ca, ob = *cb;
            va.x = fma(p.beta, oa.x, va.x); va.y = fma(p.beta, oa.y, va.y);
            vb.x = fma(p.beta, ob.x, vb.x); vb.y = fma(p.beta, ob.y, vb.y);
          }
          *ca = va;
          *cb = vb;
        }
      }
    ct = next_valid(ct + (int)gridDim.x);
    if (!bzero && ct < T) prefetch_c(ct);
  }
  cp_async_wait<0>();
}

inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

}  // namespace

bool qgemm_x_enabled() {
  const char* e = getenv("ZQ_Q8X");                 // read at every call (tests switch it)
  return e && atoi(e) != 0;
}

size_t qgemm_x_operand_doubles(int rows_max, int kmax) {
  return (size_t)8 * round_up(kmax, XBK) * round_up(rows_max, 32);
}

// C (M x N quaternions) = beta C + alpha A op(B): A stored M x K (a-part, b-part aoff behind), tb = 0: B stored K x N,
// tb = 1: B stored N x K and used as its quaternion conjugate transpose.  A8 / B8: workspaces of at least
// qgemm_x_operand_doubles(M, K) / (N, K) doubles.
void launch_qgemm_x(int tb, int M, int N, int K, double alpha, const cplx* A, size_t lda, size_t aoff, const cplx* B, size_t ldb,
                    size_t boff, double beta, cplx* C, size_t ldc, size_t coff, int lower, double* A8, double* B8, cudaStream_t st) {
  if (M <= 0 || N <= 0 || K <= 0) return;
  const int K8 = round_up(K, XBK), ldA = round_up(M, 32), ldB = round_up(N, 32);
  k_combine_a<<<dim3((ldA + 255) / 256, K8), 256, 0, st>>>(A, lda, aoff, M, K, A8, ldA, K8);
  if (tb) k_combine_b<1><<<dim3((ldB + 255) / 256, K8), 256, 0, st>>>(B, ldb, boff, N, K, B8, ldB, K8);
  else k_combine_b<0><<<dim3((ldB + 255) / 256, K8), 256, 0, st>>>(B, ldb, boff, N, K, B8, ldB, K8);
  XArgs a;
  a.M = M; a.N = N; a.K8 = K8; a.alpha = alpha; a.beta = beta;
  a.A8 = A8; a.ldA = ldA; a.B8 = B8; a.ldB = ldB; a.C = C; a.ldc = ldc; a.coff = coff; a.lower = lower;
  const size_t smem = (size_t)XST * XSTAGE * sizeof(double);
  static std::atomic<unsigned long long> attr_done{0};
  static int sms = 0;
  if (first_use_on_this_device(attr_done)) {
    cudaFuncSetAttribute(k_qgemm8x, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const int tiles_m = ldA / XBM, tiles_n = ldB / XBN;
  const long total = (long)tiles_m * tiles_n;
  const int nsm = sms > 0 ? sms : 148;
  const int grid = (int)(total < 3L * nsm ? total : 3L * nsm);
  k_qgemm8x<<<grid, XNT, smem, st>>>(a, tiles_m, tiles_n);
}

}  // namespace zq
