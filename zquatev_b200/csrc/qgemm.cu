// K4 / K6 as QUATERNION matrix products with EIGHT real products per quaternion product.
//
// The stacked complex GEMMs of zgemm.cu compute the left block column of Phi(A) Phi(B) = Phi(A B): per quaternion
// multiply-add four complex products, i.e. 16 real multiply-adds (12 with the 3M scheme the reference's zgemm3m uses,
// Makefile:2 / f77.h:80-83).  The quaternion algebra has bilinear rank 8 over the reals (Howell & Lafon 1975): with
//   q = q0 + q1 i + q2 j + q3 k,   a = (a1..a4), b = (b1..b4) the components of the two factors,
//   p1 = (a4+a2)(b2+b3)   p2 = (a1-a3)(b1+b4)   p3 = (a1+a3)(b1-b4)   p4 = (a4-a2)(b2-b3)
//   p5 = (a4-a3)(b3-b4)   p6 = (a2+a1)(b2+b1)   p7 = (a1-a2)(b3+b4)   p8 = (a4+a3)(b1-b2)
//   s  = (p1+p2+p3+p4)/2
//   c1 = s - p1 + p5      c2 = s - (p1+p2+p3) + p6      c3 = s - p2 + p7      c4 = s - p3 + p8
// every product has the a-combination on the left and the b-combination on the right, so the identity holds for
// MATRICES of components: a quaternion GEMM is eight real GEMMs whose operands are +/- sums of the four component
// planes, plus an O(MN) recombination.  Here the eight real products run on the FP64 tensor path (DMMA m8n8k4) with
// eight accumulator sets per thread; the component sums are formed in registers from the (a, b) complex-pair
// fragments, and the recombination is the epilogue.  8 instead of 12 DMMAs per quaternion multiply-add.  Accuracy is
// that of the 3M scheme (normwise, not componentwise): the test suite holds a numpy restatement (quat_kernels.qgemm8)
// and checks the back-transformation built on it against the stacked complex form.
//
// Operand layout: a quaternion matrix Q = Qa + j Qb is a pair of complex column-major arrays, the b-part `off`
// elements behind the a-part (the solver's own layout: D/E, Xa/Xb, stacked panels).  Components: q0 = Re Qa,
// q1 = Im Qa, q2 = Re Qb, q3 = -Im Qb.
//   TA = 0: A(m,k) = SA[m + k lda]          TA = 1: A = SA^H, SA stored K x M:  A(m,k) = conj_q(SA[k + m lda])
//   TB = 0: B(k,n) = SB[k + n ldb]          TB = 1: B = SB^H, SB stored N x K:  B(k,n) = conj_q(SB[n + k ldb])
//   C(m,n) <- beta C + alpha A B   (alpha, beta real);  lower != 0: only entries with m >= n are touched.
// Tiling: CTA = 4 warps (2 x 2), warp tile 16 x 16 quaternions (2 x 2 DMMA fragments x 8 planes = 64 accumulator
// doubles), CTA tile 32 x 32, BK = 8, 3-stage cp.async pipeline + the C tile staged by cp.async, two CTAs per SM.
#include "kernels.h"
#include "gemm_tiles.cuh"

namespace zq {
namespace {

constexpr int QBM = 32, QBN = 32, QNT = 128;

struct QArgs {
  int M, N, K;
  double alpha, beta;
  const cplx* A; size_t lda, aoff;
  const cplx* B; size_t ldb, boff;
  cplx* C; size_t ldc, coff;
  int lower;
  int cb0, cbs;           // column-block addressing (multi-GPU trailing update): 64-column block index = cb0 + i * cbs
  size_t sA, sB, sC;      // batch strides (elements)
  SplitK sk;              // chunks > 0: split-K, part z writes C + z * sC (alpha = 1, beta = 0 expected)
};

// combos_a / combos_b (the eight left / right combinations of the four components) live in gemm_tiles.cuh

constexpr int QCLD = QBM + 1;          // column stride of the staged C tile (odd: the epilogue's quarter-warp reads hit 8 bank groups)
constexpr int QCEL = 2 * QBN * QCLD;   // complex elements of the staged C tile (a-part, b-part)

// CPRE: the C tile (read-modify-write, beta != 0) is fetched into shared memory by cp.async at kernel start, so the
// epilogue of a short-K product (K6 update: K = 64) does not wait on global loads.
template <int TA, int TB, int QBK, int QST, int MINB, bool CPRE>
__global__ void __launch_bounds__(QNT, MINB) k_qgemm8(QArgs p) {
  using TileA = OpTile<QBM, TA == 1, QBK>;
  using TileB = OpTile<QBN, TB == 0, QBK>;
  constexpr int STAGE = 2 * TileA::ELEMS + 2 * TileB::ELEMS;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cplx* smem = reinterpret_cast<cplx*>(smem_raw);

  const int r0 = blockIdx.x * QBM;
  const int c0 = (p.cb0 + (int)(blockIdx.y >> 1) * p.cbs) * 64 + (int)(blockIdx.y & 1) * QBN;   // an owned 64-column block = two tiles
  if (c0 >= p.N || (p.lower && r0 + QBM - 1 < c0)) return;
  const cplx* A = p.A;
  const cplx* B = p.B;
  int K = p.K;
  if (p.sk.chunks > 0) {
    const int k0 = (blockIdx.z % p.sk.chunks) * p.sk.kc;
    A += TA ? (size_t)k0 : (size_t)k0 * p.lda;
    B += TB ? (size_t)k0 * p.ldb : (size_t)k0;
    K = min(p.sk.kc, K - k0);
  } else {
    A += (size_t)blockIdx.z * p.sA;
    B += (size_t)blockIdx.z * p.sB;
  }
  cplx* C = p.C + (size_t)blockIdx.z * p.sC;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = (warp & 1) * 16, wn = (warp >> 1) * 16;
  const int g = lane >> 2, q = lane & 3;
  const bool bzero = (p.beta == 0.0);

  cplx* sC = smem + (size_t)QST * STAGE;
  if (CPRE && !bzero) {
    for (int e = tid; e < 2 * QBN * QBM; e += QNT) {
      const int r = e % QBM, c = (e / QBM) % QBN, half = e / (QBM * QBN);
      const bool ok = (r0 + r < p.M) && (c0 + c < p.N);
      cp_async16(sC + (half * QBN + c) * QCLD + r, ok ? C + (size_t)half * p.coff + (size_t)(r0 + r) + (size_t)(c0 + c) * p.ldc : C, ok);
    }
    cp_async_commit();
  }
  if (!CPRE && !bzero) {   // warm L2 with the C tile this CTA will read-modify-write in the epilogue
    for (int e = tid; e < QBN * (QBM / 8) * 2; e += QNT) {
      const int half = e / (QBN * (QBM / 8)), f = e % (QBN * (QBM / 8));
      const int c = c0 + f / (QBM / 8), r = r0 + (f % (QBM / 8)) * 8;
      if (c < p.N && r < p.M) asm volatile("prefetch.global.L2 [%0];\n" ::"l"(C + (size_t)half * p.coff + (size_t)r + (size_t)c * p.ldc));
    }
  }

  double acc[8][2][2][2];
#pragma unroll
  for (int e = 0; e < 8; ++e)
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) acc[e][i][j][0] = acc[e][i][j][1] = 0.0;

  const int nk = (K + QBK - 1) / QBK;
  TileLoader<QBM, TA == 1, QBK, QNT> ldAa, ldAb;
  TileLoader<QBN, TB == 0, QBK, QNT> ldBa, ldBb;
  ldAa.init(A, p.lda, r0, p.M, tid);
  ldAb.init(A + p.aoff, p.lda, r0, p.M, tid);
  ldBa.init(B, p.ldb, c0, p.N, tid);
  ldBb.init(B + p.boff, p.ldb, c0, p.N, tid);
  auto issue = [&](int kt) {
    if (kt < nk) {
      cplx* s = smem + (size_t)(kt % QST) * STAGE;
      ldAa.issue(s, kt * QBK, K);
      ldAb.issue(s + TileA::ELEMS, kt * QBK, K);
      ldBa.issue(s + 2 * TileA::ELEMS, kt * QBK, K);
      ldBb.issue(s + 2 * TileA::ELEMS + TileB::ELEMS, kt * QBK, K);
    }
    cp_async_commit();
  };
#pragma unroll
  for (int s = 0; s < QST - 1; ++s) issue(s);

  for (int kt = 0; kt < nk; ++kt) {
    cp_async_wait<QST - 2>();
    __syncthreads();
    issue(kt + QST - 1);
    const cplx* sAa = smem + (size_t)(kt % QST) * STAGE;
    const cplx* sAb = sAa + TileA::ELEMS;
    const cplx* sBa = sAb + TileA::ELEMS;
    const cplx* sBb = sBa + TileB::ELEMS;
#pragma unroll
    for (int k4 = 0; k4 < QBK; k4 += 4) {
      double al[2][8], be[2][8];
#pragma unroll
      for (int i = 0; i < 2; ++i)
        combos_a<TA == 1>(TileA::frag(sAa, wm + 8 * i + g, k4 + q), TileA::frag(sAb, wm + 8 * i + g, k4 + q), al[i]);
#pragma unroll
      for (int j = 0; j < 2; ++j)
        combos_b<TB == 1>(TileB::frag(sBa, wn + 8 * j + g, k4 + q), TileB::frag(sBb, wn + 8 * j + g, k4 + q), be[j]);
#pragma unroll
      for (int e = 0; e < 8; ++e)
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int j = 0; j < 2; ++j) dmma(acc[e][i][j][0], acc[e][i][j][1], al[i][e], be[j][e]);
    }
  }
  cp_async_wait<0>();

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
        const double s = 0.5 * (s123 + p4);
        const double q0 = (s - p1) + acc[4][i][j][h];
        const double q1 = (s - s123) + acc[5][i][j][h];
        const double q2 = (s - p2) + acc[6][i][j][h];
        const double q3 = (s - p3) + acc[7][i][j][h];
        cplx va = cmake(p.alpha * q0, p.alpha * q1);
        cplx vb = cmake(p.alpha * q2, -p.alpha * q3);
        cplx* ca = C + (size_t)r + (size_t)c * p.ldc;
        cplx* cb = ca + p.coff;
        if (!bzero) {
          const cplx oa = CPRE ? sC[(c - c0) * QCLD + (r - r0)] : *ca;
          const cplx ob = CPRE ? sC[(QBN + c - c0) * QCLD + (r - r0)] : *cb;
          va.x = fma(p.beta, oa.x, va.x); va.y = fma(p.beta, oa.y, va.y);
          vb.x = fma(p.beta, ob.x, vb.x); vb.y = fma(p.beta, ob.y, vb.y);
        }
        *ca = va;
        *cb = vb;
      }
    }
}

template <int TA, int TB, int QBK, int QST, int MINB, bool CPRE>
void launch_q_cfg(const QArgs& a, int gz, int ncb, cudaStream_t st) {
  using TileA = OpTile<QBM, TA == 1, QBK>;
  using TileB = OpTile<QBN, TB == 0, QBK>;
  const size_t smem = ((size_t)QST * (2 * TileA::ELEMS + 2 * TileB::ELEMS) + (CPRE ? QCEL : 0)) * sizeof(cplx);
  static std::atomic<unsigned long long> attr_done{0};
  if (first_use_on_this_device(attr_done))
    cudaFuncSetAttribute(k_qgemm8<TA, TB, QBK, QST, MINB, CPRE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  dim3 g((a.M + QBM - 1) / QBM, ncb >= 0 ? 2 * ncb : 2 * ((a.N + 63) / 64), gz);
  if (g.y == 0) return;
  k_qgemm8<TA, TB, QBK, QST, MINB, CPRE><<<g, QNT, smem, st>>>(a);
}

// ---------------------------------------------------------------------------------------------
// Persistent variant for the short-K products (K4: K = 128, K6 update: K = 64; no split-K): 2 CTAs per SM loop over the
// output tiles (round-robin), and the cp.async operand pipeline keeps running ACROSS tile boundaries -- while a tile's
// epilogue recombines and stores, the first stages of the next tile are already in flight, and the next C tile is
// fetched into shared memory as soon as the epilogue has read the current one.  Removes the per-tile prologue / epilogue
// bubbles that leave the shared FP64 pipe 64-74 % busy in the one-tile-per-CTA kernel (profiles/r02_ncu_qgemm.md).
// ---------------------------------------------------------------------------------------------
template <int TA, int TB, int QBK, int QST>
__global__ void __launch_bounds__(QNT, 2) k_qgemm8p(QArgs p, int tiles_m, int tiles_n) {
  using TileA = OpTile<QBM, TA == 1, QBK>;
  using TileB = OpTile<QBN, TB == 0, QBK>;
  constexpr int STAGE = 2 * TileA::ELEMS + 2 * TileB::ELEMS;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cplx* smem = reinterpret_cast<cplx*>(smem_raw);
  cplx* sC = smem + (size_t)QST * STAGE;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = (warp & 1) * 16, wn = (warp >> 1) * 16;
  const int g = lane >> 2, q = lane & 3;
  const bool bzero = (p.beta == 0.0);
  const int K = p.K, nk = (K + QBK - 1) / QBK;
  const int T = tiles_m * tiles_n;
  // tile t -> (r0, c0); false: nothing to do there (beyond N, or above the diagonal of a `lower` product)
  auto tile_rc = [&](int t, int& r0, int& c0) -> bool {
    const int tm = t % tiles_m, tn = t / tiles_m;
    r0 = tm * QBM;
    c0 = (p.cb0 + (tn >> 1) * p.cbs) * 64 + (tn & 1) * QBN;
    return c0 < p.N && !(p.lower && r0 + QBM - 1 < c0);
  };
  auto next_valid = [&](int t) -> int {
    int r0, c0;
    while (t < T && !tile_rc(t, r0, c0)) t += gridDim.x;
    return t;
  };
  auto fetch_c = [&](int t) {          // cp.async of the C tile of tile t (joins the next committed group)
    int r0, c0;
    tile_rc(t, r0, c0);
    for (int e = tid; e < 2 * QBN * QBM; e += QNT) {
      const int r = e % QBM, c = (e / QBM) % QBN, half = e / (QBM * QBN);
      const bool ok = (r0 + r < p.M) && (c0 + c < p.N);
      cp_async16(sC + (half * QBN + c) * QCLD + r, ok ? p.C + (size_t)half * p.coff + (size_t)(r0 + r) + (size_t)(c0 + c) * p.ldc : p.C, ok);
    }
  };

  // loader cursor (runs QST - 1 stages ahead of the consumer, across tile boundaries)
  TileLoader<QBM, TA == 1, QBK, QNT> ldAa, ldAb;
  TileLoader<QBN, TB == 0, QBK, QNT> ldBa, ldBb;
  int lt = next_valid((int)blockIdx.x), lkt = 0;
  unsigned gl = 0;                     // stages issued so far (slot = gl % QST)
  auto loader_init = [&]() {
    int r0, c0;
    tile_rc(lt, r0, c0);
    ldAa.init(p.A, p.lda, r0, p.M, tid);
    ldAb.init(p.A + p.aoff, p.lda, r0, p.M, tid);
    ldBa.init(p.B, p.ldb, c0, p.N, tid);
    ldBb.init(p.B + p.boff, p.ldb, c0, p.N, tid);
  };
  auto issue_next = [&]() {
    if (lt < T) {
      cplx* s = smem + (size_t)(gl % QST) * STAGE;
      ldAa.issue(s, lkt * QBK, K);
      ldAb.issue(s + TileA::ELEMS, lkt * QBK, K);
      ldBa.issue(s + 2 * TileA::ELEMS, lkt * QBK, K);
      ldBb.issue(s + 2 * TileA::ELEMS + TileB::ELEMS, lkt * QBK, K);
      if (++lkt == nk) {
        lkt = 0;
        lt = next_valid(lt + (int)gridDim.x);
        if (lt < T) loader_init();
      }
    }
    ++gl;
    cp_async_commit();
  };

  int ct = lt;                         // consumer cursor
  if (ct >= T) return;
  loader_init();
  if (!bzero) fetch_c(ct);
#pragma unroll
  for (int s = 0; s < QST - 1; ++s) issue_next();
  unsigned gu = 0;                     // stages consumed so far

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
      cp_async_wait<QST - 2>();
      __syncthreads();
      issue_next();
      const cplx* sAa = smem + (size_t)(gu % QST) * STAGE;
      const cplx* sAb = sAa + TileA::ELEMS;
      const cplx* sBa = sAb + TileA::ELEMS;
      const cplx* sBb = sBa + TileB::ELEMS;
#pragma unroll
      for (int k4 = 0; k4 < QBK; k4 += 4) {
        double al[2][8], be[2][8];
#pragma unroll
        for (int i = 0; i < 2; ++i)
          combos_a<TA == 1>(TileA::frag(sAa, wm + 8 * i + g, k4 + q), TileA::frag(sAb, wm + 8 * i + g, k4 + q), al[i]);
#pragma unroll
        for (int j = 0; j < 2; ++j)
          combos_b<TB == 1>(TileB::frag(sBa, wn + 8 * j + g, k4 + q), TileB::frag(sBb, wn + 8 * j + g, k4 + q), be[j]);
#pragma unroll
        for (int e = 0; e < 8; ++e)
#pragma unroll
          for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 2; ++j) dmma(acc[e][i][j][0], acc[e][i][j][1], al[i][e], be[j][e]);
      }
    }
    if (!bzero && nk < QST) {          // very short K: the C tile's group may still be in flight
      cp_async_wait<0>();
      __syncthreads();
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
            const cplx oa = sC[(c - c0) * QCLD + (r - r0)];
            const cplx ob = sC[(QBN + c - c0) * QCLD + (r - r0)];
            va.x = fma(p.beta, oa.x, va.x); va.y = fma(p.beta, oa.y, va.y);
            vb.x = fma(p.beta, ob.x, vb.x); vb.y = fma(p.beta, ob.y, vb.y);
          }
          *ca = va;
          *cb = vb;
        }
      }
    ct = next_valid(ct + (int)gridDim.x);
    if (!bzero && ct < T) {
      __syncthreads();                 // every thread has read the staged C tile
      fetch_c(ct);                     // joins the group committed by the next issue_next()
    }
  }
  cp_async_wait<0>();
}

template <int TA, int TB, int QBK, int QST>
void launch_q_persistent(const QArgs& a, int ncb, cudaStream_t st) {
  using TileA = OpTile<QBM, TA == 1, QBK>;
  using TileB = OpTile<QBN, TB == 0, QBK>;
  const size_t smem = ((size_t)QST * (2 * TileA::ELEMS + 2 * TileB::ELEMS) + QCEL) * sizeof(cplx);
  static std::atomic<unsigned long long> attr_done{0};
  static int sms = 0;
  if (first_use_on_this_device(attr_done)) {
    cudaFuncSetAttribute(k_qgemm8p<TA, TB, QBK, QST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const int tiles_m = (a.M + QBM - 1) / QBM, tiles_n = ncb >= 0 ? 2 * ncb : 2 * ((a.N + 63) / 64);
  const long total = (long)tiles_m * tiles_n;
  if (total <= 0) return;
  const int nsm = sms > 0 ? sms : 148;
  const int grid = (int)(total < 2L * nsm ? total : 2L * nsm);
  k_qgemm8p<TA, TB, QBK, QST><<<grid, QNT, smem, st>>>(a, tiles_m, tiles_n);
}

// ZQ_Q8_CFG (development knob; measured on the 2n = 32768 shapes, profiles/r02_gemm_probe_q8_cfgs.jsonl): 3 (default) BK 8 x
// 3 stages + the C tile staged in shared memory, 2 CTAs/SM (back-transformation 3074 ms); 0: BK 8 x 4 stages, C read in the
// epilogue (3226 ms); 1: BK 8 x 3 stages, 3 CTAs/SM under a 168-register cap (3239 ms: the spills cost what the third CTA
// buys); 2: BK 16 x 2 stages; 4: as 3 with 2 stages and 3 CTAs/SM; 5: BK 16 x 2 stages + staged C tile
template <int TA, int TB>
void launch_q(const QArgs& a, int gz, int ncb, cudaStream_t st) {
  static const int cfg = [] { const char* e = getenv("ZQ_Q8_CFG"); return e ? atoi(e) : 3; }();
  // ZQ_Q8_PERSIST (default 1): persistent CTAs for the plain (one batch, no split-K) products -- K6 update 44.7 -> 47.5, merged
  // two-panel update 49.7 -> 51.6 canonical TFLOP/s (profiles/r02_gemm_probe_pairs.jsonl)
  static const int persist = [] { const char* e = getenv("ZQ_Q8_PERSIST"); return e ? atoi(e) : 1; }();
  if (persist && gz == 1 && a.sk.chunks == 0) {
    launch_q_persistent<TA, TB, 8, 3>(a, ncb, st);
    return;
  }
  if (cfg == 1) launch_q_cfg<TA, TB, 8, 3, 3, false>(a, gz, ncb, st);
  else if (cfg == 2) launch_q_cfg<TA, TB, 16, 2, 2, false>(a, gz, ncb, st);
  else if (cfg == 0) launch_q_cfg<TA, TB, 8, 4, 2, false>(a, gz, ncb, st);
  else if (cfg == 4) launch_q_cfg<TA, TB, 8, 2, 3, true>(a, gz, ncb, st);
  else if (cfg == 5) launch_q_cfg<TA, TB, 16, 2, 2, true>(a, gz, ncb, st);
  else launch_q_cfg<TA, TB, 8, 3, 2, true>(a, gz, ncb, st);
}

}  // namespace

void launch_qgemm(int ta, int tb, int M, int N, int K, double alpha, const cplx* A, size_t lda, size_t aoff, const cplx* B,
                  size_t ldb, size_t boff, double beta, cplx* C, size_t ldc, size_t coff, int lower, int batch, size_t sA,
                  size_t sB, size_t sC, const SplitK* sk, cudaStream_t st, int cb0, int cbs, int ncb) {
  if (M <= 0 || N <= 0 || batch <= 0) return;
  QArgs a;
  a.M = M; a.N = N; a.K = K; a.alpha = alpha; a.beta = beta;
  a.A = A; a.lda = lda; a.aoff = aoff; a.B = B; a.ldb = ldb; a.boff = boff; a.C = C; a.ldc = ldc; a.coff = coff;
  a.lower = lower; a.sA = sA; a.sB = sB; a.sC = sC; a.cb0 = cb0; a.cbs = cbs;
  a.sk = sk ? *sk : SplitK{};
  const int gz = (sk && sk->chunks > 0) ? sk->chunks : batch;
  if (ta == 0 && tb == 0) launch_q<0, 0>(a, gz, ncb, st);
  else if (ta == 0 && tb == 1) launch_q<0, 1>(a, gz, ncb, st);
  else if (ta == 1 && tb == 0) launch_q<1, 0>(a, gz, ncb, st);
  else launch_q<1, 1>(a, gz, ncb, st);
}

}  // namespace zq
