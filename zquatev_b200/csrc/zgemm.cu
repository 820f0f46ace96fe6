// K4 / K6: FP64 complex GEMM  C = alpha * op(A) * op(B) + beta * C  (column-major) on the FP64
// tensor path (DMMA, mma.sync.m8n8k4.f64).
//
// Serves (a) the trailing rank-2k update of the tridiagonalisation -- the ONE stacked
// contraction [D;E] -= L R^H that replaces the eight zgemm3m calls of the reference
// (blocked.cc:424-474) -- with `lower` restricting work to the lower triangles, and (b) the
// compact-WY back-transformation Y = Phi(V)^H X, X -= Phi(V) (T Y) that replaces the
// reference's forward accumulation of Q (blocked.cc:477-544) and final zgemm3m pair
// (zquatev.cc:87-90).
//
// Why DMMA: on B200 the FP64 FMA pipe and the FP64 tensor path have the same measured peak
// (tools/fp64_peak: 36.9 vs 37.1 TFLOP/s) but one DMMA replaces 8 DFMA per lane, so the issue
// slots, register-file ports and shared-memory bandwidth that a register-tiled DFMA kernel needs
// (v1 of this file reached 15-20 TFLOP/s) are freed.  A complex product is four real DMMAs on
// (re, im) fragments; conjugation is a sign flip of the imaginary fragment in registers.
//
// Tiling (defaults; measured alternatives are listed where they are selected): the 4-product kernel k_zgemm_mma uses a
// 64 x 64 CTA tile (4 warps, warp tile 32 x 32), the 3M kernel k_zgemm_3m a 64 x 32 CTA tile (4 warps, warp tile
// 32 x 16, three accumulator sets); both BK = 8 with a 3-stage cp.async pipeline straight from global to (padded,
// bank-conflict-free) shared memory -- operands are stored interleaved complex exactly as in global memory, in
// whichever of the two orientations (contiguous along the tile dimension, or contiguous along k) the caller's op
// needs.  The larger mma.sync .f64 shapes of the PTX ISA (m16n8k4/k8/k16) were checked: ptxas for sm_100a expands
// each of them into DMMA.8x8x4 sequences (cuobjdump of tools/fp64_peak.cu: 1160 DMMA.8x8x4, no other DMMA shape), and
// their fragment layout is the m8n8k4 layout repeated, so neither instruction count in SASS nor shared-memory loads
// change -- m8n8k4 is the native FP64 tensor instruction of this chip.
#include "kernels.h"
#include "gemm_tiles.cuh"

namespace zq {
namespace {


thread_local int g_allow_3m = 1;   // set per solve by the calling thread (launches happen on that thread): small problems use the conventional product

// TA/TB: 0 = operand used as stored, 1 = conjugate transpose.
//   A as stored (TA=0) is M x K (contiguous along m)  -> KCONT = false
//   A^H       (TA=1) is stored K x M (contiguous along k) -> KCONT = true, conj
//   B as stored (TB=0) is K x N (contiguous along k)  -> KCONT = true
//   B^H       (TB=1) is stored N x K (contiguous along n) -> KCONT = false, conj
template <int BM, int BN, int BK, int STAGES, int TA, int TB>
__global__ void __launch_bounds__((BM / 32) * (BN / 32) * 32)
k_zgemm_mma(int M, int N, int K, cplx alpha, const cplx* __restrict__ A, size_t lda, const cplx* __restrict__ B,
            size_t ldb, cplx beta, cplx* __restrict__ C, size_t ldc, int lower, size_t sA, size_t sB, size_t sC, int cb0, int cbs,
            SplitK sk) {
  constexpr int NTHREADS = (BM / 32) * (BN / 32) * 32;
  using TileA = OpTile<BM, TA == 1, BK>;
  using TileB = OpTile<BN, TB == 0, BK>;
  constexpr int STAGE_ELEMS = TileA::ELEMS + TileB::ELEMS;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cplx* smem = reinterpret_cast<cplx*>(smem_raw);

  const int r0 = blockIdx.x * BM, c0 = (cb0 + blockIdx.y * cbs) * BN;
  if (lower && r0 + BM - 1 < c0) return;
  if (sk.chunks > 0) {
    // split-K: grid.z = segments x chunks; part z covers k in [c*kc, c*kc + kc) of segment seg and owns its own C
    const int seg = blockIdx.z / sk.chunks, k0 = (blockIdx.z % sk.chunks) * sk.kc;
    A += (size_t)seg * sk.segA + (TA ? (size_t)k0 : (size_t)k0 * lda);
    B += (size_t)seg * sk.segB + (TB ? (size_t)k0 * ldb : (size_t)k0);
    K = min(sk.kc, K - k0);
  } else {
    A += (size_t)blockIdx.z * sA;
    B += (size_t)blockIdx.z * sB;
  }
  C += (size_t)blockIdx.z * sC;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = (warp % (BM / 32)) * 32, wn = (warp / (BM / 32)) * 32;
  const int g = lane >> 2, q = lane & 3;

  // warm L2 with the C tile this CTA will read-modify-write in the epilogue
  const bool bzero = (beta.x == 0.0 && beta.y == 0.0);
  if (!bzero) {
    for (int e = tid; e < BN * (BM / 8); e += NTHREADS) {
      const int c = c0 + e / (BM / 8), r = r0 + (e % (BM / 8)) * 8;
      if (c < N && r < M) asm volatile("prefetch.global.L2 [%0];\n" ::"l"(C + (size_t)r + (size_t)c * ldc));
    }
  }

  double cre[4][4][2], cim[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { cre[i][j][0] = cre[i][j][1] = 0.0; cim[i][j][0] = cim[i][j][1] = 0.0; }

  const int nk = (K + BK - 1) / BK;
  TileLoader<BM, TA == 1, BK, NTHREADS> ldA;
  TileLoader<BN, TB == 0, BK, NTHREADS> ldB;
  ldA.init(A, lda, r0, M, tid);
  ldB.init(B, ldb, c0, N, tid);
  auto issue = [&](int kt) {
    if (kt < nk) {
      cplx* sa = smem + (size_t)(kt % STAGES) * STAGE_ELEMS;
      ldA.issue(sa, kt * BK, K);
      ldB.issue(sa + TileA::ELEMS, kt * BK, K);
    }
    cp_async_commit();
  };
#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) issue(s);

  for (int kt = 0; kt < nk; ++kt) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    issue(kt + STAGES - 1);
    const cplx* sa = smem + (size_t)(kt % STAGES) * STAGE_ELEMS;
    const cplx* sb = sa + TileA::ELEMS;
#pragma unroll
    for (int k4 = 0; k4 < BK; k4 += 4) {
      double ar[4], ai[4], nai[4], br[4], bi[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const cplx a = TileA::frag(sa, wm + 8 * i + g, k4 + q);
        ar[i] = a.x;
        ai[i] = TA ? -a.y : a.y;
        nai[i] = -ai[i];
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const cplx b = TileB::frag(sb, wn + 8 * j + g, k4 + q);
        br[j] = b.x;
        bi[j] = TB ? -b.y : b.y;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          dmma(cre[i][j][0], cre[i][j][1], ar[i], br[j]);
          dmma(cim[i][j][0], cim[i][j][1], ar[i], bi[j]);
        }
      // second half in a separate sweep: 32 independent DMMAs sit between the two updates of an accumulator
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          dmma(cre[i][j][0], cre[i][j][1], nai[i], bi[j]);
          dmma(cim[i][j][0], cim[i][j][1], ai[i], br[j]);
        }
    }
  }
  cp_async_wait<0>();

  // epilogue: lane holds rows wm+8i+g, columns wn+8j+2q+{0,1}
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int c = c0 + wn + 8 * j + 2 * q + h;
      if (c >= N) continue;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = r0 + wm + 8 * i + g;
        if (r >= M || (lower && r < c)) continue;
        cplx v = cmul(alpha, cmake(cre[i][j][h], cim[i][j][h]));
        cplx* cp = C + (size_t)r + (size_t)c * ldc;
        if (!bzero) { const cplx o = *cp; cfma(v, beta, o); }
        *cp = v;
      }
    }
}

// ---------------------------------------------------------------------------------------------
// 3M variant: a complex product from THREE real DMMAs instead of four,
//   T1 = Ar Br,  T2 = Ai Bi,  T3 = (Ar + Ai)(Br + Bi);   Re = T1 - T2,  Im = T3 - T1 - T2,
// i.e. 25 % fewer tensor-pipe cycles for the same result -- the same trade the reference makes by
// calling zgemm3m everywhere (Makefile:2, f77.h:80-83).  The fragment sums are formed in registers;
// three accumulator sets are kept, so the warp tile is 32 x 16 (CTA 64 x 32, 4 warps, 3 CTAs/SM).
// Column-block addressing (multi-GPU trailing update): an owned 64-column block = two 32-column tiles.
// ---------------------------------------------------------------------------------------------
template <int BK, int STAGES, int TA, int TB>
__global__ void __launch_bounds__(128)
k_zgemm_3m(int M, int N, int K, cplx alpha, const cplx* __restrict__ A, size_t lda, const cplx* __restrict__ B,
           size_t ldb, cplx beta, cplx* __restrict__ C, size_t ldc, int lower, size_t sA, size_t sB, size_t sC, int cb0, int cbs,
           SplitK sk) {
  constexpr int BM = 64, BN = 32, NTHREADS = 128;
  using TileA = OpTile<BM, TA == 1, BK>;
  using TileB = OpTile<BN, TB == 0, BK>;
  constexpr int STAGE_ELEMS = TileA::ELEMS + TileB::ELEMS;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cplx* smem = reinterpret_cast<cplx*>(smem_raw);

  const int r0 = blockIdx.x * BM;
  const int c0 = (cb0 + (blockIdx.y >> 1) * cbs) * 64 + (blockIdx.y & 1) * BN;
  if (c0 >= N || (lower && r0 + BM - 1 < c0)) return;
  if (sk.chunks > 0) {
    // split-K: grid.z = segments x chunks; part z covers k in [c*kc, c*kc + kc) of segment seg and owns its own C
    const int seg = blockIdx.z / sk.chunks, k0 = (blockIdx.z % sk.chunks) * sk.kc;
    A += (size_t)seg * sk.segA + (TA ? (size_t)k0 : (size_t)k0 * lda);
    B += (size_t)seg * sk.segB + (TB ? (size_t)k0 * ldb : (size_t)k0);
    K = min(sk.kc, K - k0);
  } else {
    A += (size_t)blockIdx.z * sA;
    B += (size_t)blockIdx.z * sB;
  }
  C += (size_t)blockIdx.z * sC;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = (warp & 1) * 32, wn = (warp >> 1) * 16;
  const int g = lane >> 2, q = lane & 3;

  const bool bzero = (beta.x == 0.0 && beta.y == 0.0);
  if (!bzero) {
    for (int e = tid; e < BN * (BM / 8); e += NTHREADS) {
      const int c = c0 + e / (BM / 8), r = r0 + (e % (BM / 8)) * 8;
      if (c < N && r < M) asm volatile("prefetch.global.L2 [%0];\n" ::"l"(C + (size_t)r + (size_t)c * ldc));
    }
  }

  double t1[4][2][2], t2[4][2][2], t3[4][2][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int h = 0; h < 2; ++h) t1[i][j][h] = t2[i][j][h] = t3[i][j][h] = 0.0;

  const int nk = (K + BK - 1) / BK;
  TileLoader<BM, TA == 1, BK, NTHREADS> ldA;
  TileLoader<BN, TB == 0, BK, NTHREADS> ldB;
  ldA.init(A, lda, r0, M, tid);
  ldB.init(B, ldb, c0, N, tid);
  auto issue = [&](int kt) {
    if (kt < nk) {
      cplx* sa = smem + (size_t)(kt % STAGES) * STAGE_ELEMS;
      ldA.issue(sa, kt * BK, K);
      ldB.issue(sa + TileA::ELEMS, kt * BK, K);
    }
    cp_async_commit();
  };
#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) issue(s);

  for (int kt = 0; kt < nk; ++kt) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    issue(kt + STAGES - 1);
    const cplx* sa = smem + (size_t)(kt % STAGES) * STAGE_ELEMS;
    const cplx* sb = sa + TileA::ELEMS;
#pragma unroll
    for (int k4 = 0; k4 < BK; k4 += 4) {
      double ar[4], ai[4], as[4], br[2], bi[2], bs[2];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const cplx a = TileA::frag(sa, wm + 8 * i + g, k4 + q);
        ar[i] = a.x;
        ai[i] = TA ? -a.y : a.y;
        as[i] = ar[i] + ai[i];
      }
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const cplx b = TileB::frag(sb, wn + 8 * j + g, k4 + q);
        br[j] = b.x;
        bi[j] = TB ? -b.y : b.y;
        bs[j] = br[j] + bi[j];
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          dmma(t1[i][j][0], t1[i][j][1], ar[i], br[j]);
          dmma(t2[i][j][0], t2[i][j][1], ai[i], bi[j]);
          dmma(t3[i][j][0], t3[i][j][1], as[i], bs[j]);
        }
    }
  }
  cp_async_wait<0>();

#pragma unroll
  for (int j = 0; j < 2; ++j)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int c = c0 + wn + 8 * j + 2 * q + h;
      if (c >= N) continue;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = r0 + wm + 8 * i + g;
        if (r >= M || (lower && r < c)) continue;
        const double re = t1[i][j][h] - t2[i][j][h];
        const double im = (t3[i][j][h] - t1[i][j][h]) - t2[i][j][h];
        cplx v = cmul(alpha, cmake(re, im));
        cplx* cp = C + (size_t)r + (size_t)c * ldc;
        if (!bzero) { const cplx o = *cp; cfma(v, beta, o); }
        *cp = v;
      }
    }
}

template <int BK, int STAGES, int TA, int TB>
void launch_3m_cfg(int M, int N, int K, cplx alpha, const cplx* A, size_t lda, const cplx* B, size_t ldb, cplx beta, cplx* C,
                   size_t ldc, int lower, int batch, size_t sA, size_t sB, size_t sC, int cb0, int cbs, int ncb, SplitK sk,
                   cudaStream_t st) {
  using TileA = OpTile<64, TA == 1, BK>;
  using TileB = OpTile<32, TB == 0, BK>;
  const size_t smem = (size_t)STAGES * (TileA::ELEMS + TileB::ELEMS) * sizeof(cplx);
  static std::atomic<unsigned long long> attr_done{0};
  if (first_use_on_this_device(attr_done)) {
    cudaFuncSetAttribute(k_zgemm_3m<BK, STAGES, TA, TB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  }
  dim3 g((M + 63) / 64, ncb >= 0 ? 2 * ncb : 2 * ((N + 63) / 64), batch);
  if (g.y == 0) return;
  k_zgemm_3m<BK, STAGES, TA, TB><<<g, 128, smem, st>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, lower, sA, sB, sC, cb0, cbs, sk);
}

// ZQ_3M_CFG (development knob): pipeline shape of the 3M kernel -- 0 (default): BK 8 x 3 stages; 1: BK 16 x 2; 2: BK 8 x 4;
// 3: BK 16 x 3 (2 CTAs per SM)
template <int TA, int TB>
void launch_3m(int M, int N, int K, cplx alpha, const cplx* A, size_t lda, const cplx* B, size_t ldb, cplx beta, cplx* C,
               size_t ldc, int lower, int batch, size_t sA, size_t sB, size_t sC, int cb0, int cbs, int ncb, SplitK sk,
               cudaStream_t st) {
  static const int cfg = [] { const char* e = getenv("ZQ_3M_CFG"); return e ? atoi(e) : 0; }();
  if (cfg == 1) launch_3m_cfg<16, 2, TA, TB>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, lower, batch, sA, sB, sC, cb0, cbs, ncb, sk, st);
  else if (cfg == 2) launch_3m_cfg<8, 4, TA, TB>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, lower, batch, sA, sB, sC, cb0, cbs, ncb, sk, st);
  else if (cfg == 3) launch_3m_cfg<16, 3, TA, TB>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, lower, batch, sA, sB, sC, cb0, cbs, ncb, sk, st);
  else launch_3m_cfg<8, 3, TA, TB>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, lower, batch, sA, sB, sC, cb0, cbs, ncb, sk, st);
}

template <int BM, int BN, int BK, int STAGES, int TA, int TB>
void launch_cfg(int M, int N, int K, cplx alpha, const cplx* A, size_t lda, const cplx* B, size_t ldb, cplx beta, cplx* C,
                size_t ldc, int lower, int batch, size_t sA, size_t sB, size_t sC, int cb0, int cbs, int ncb, SplitK sk,
                cudaStream_t st) {
  using TileA = OpTile<BM, TA == 1, BK>;
  using TileB = OpTile<BN, TB == 0, BK>;
  constexpr int NTHREADS = (BM / 32) * (BN / 32) * 32;
  const size_t smem = (size_t)STAGES * (TileA::ELEMS + TileB::ELEMS) * sizeof(cplx);
  static std::atomic<unsigned long long> attr_done{0};
  if (first_use_on_this_device(attr_done)) {
    cudaFuncSetAttribute(k_zgemm_mma<BM, BN, BK, STAGES, TA, TB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  }
  dim3 g((M + BM - 1) / BM, ncb >= 0 ? ncb : (N + BN - 1) / BN, batch);
  if (g.y == 0) return;
  k_zgemm_mma<BM, BN, BK, STAGES, TA, TB><<<g, NTHREADS, smem, st>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, lower, sA, sB, sC, cb0, cbs, sk);
}

template <int TA, int TB>
void launch_t(int M, int N, int K, cplx alpha, const cplx* A, size_t lda, const cplx* B, size_t ldb, cplx beta, cplx* C,
              size_t ldc, int lower, int batch, size_t sA, size_t sB, size_t sC, int cb0, int cbs, int ncb, SplitK sk,
              cudaStream_t st) {
  // ZQ_GEMM_CFG (development knob): 0 auto, 1: 64x128 BK16 x3, 2: 64x64 BK16 x2 (2 CTAs/SM), 3: 64x64 BK8 x3
  static const int cfg_env = [] { const char* e = getenv("ZQ_GEMM_CFG"); return e ? atoi(e) : 0; }();
  // ZQ_GEMM_3M: 1 (default) = three-multiplication complex product (k_zgemm_3m), 0 = conventional four
  static const int use_3m = [] { const char* e = getenv("ZQ_GEMM_3M"); return e ? atoi(e) : 1; }();
  if (use_3m && g_allow_3m && cfg_env == 0) {
    launch_3m<TA, TB>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, lower, batch, sA, sB, sC, cb0, cbs, ncb, sk, st);
    return;
  }
  int cfg = cfg_env;
  if (ncb >= 0) cfg = 3;   // column-block addressing assumes BN = 64
  if (cfg == 0) {
    cfg = 3;   // measured best on every shape of the solver (profiles/r01_gemm_cfg1.jsonl vs r01_gemm_cfg3.jsonl)
  }
  if (cfg == 1)
    launch_cfg<64, 128, 16, 3, TA, TB>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, lower, batch, sA, sB, sC, cb0, cbs, ncb, sk, st);
  else if (cfg == 2)
    launch_cfg<64, 64, 16, 2, TA, TB>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, lower, batch, sA, sB, sC, cb0, cbs, ncb, sk, st);
  else
    launch_cfg<64, 64, 8, 3, TA, TB>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, lower, batch, sA, sB, sC, cb0, cbs, ncb, sk, st);
}

}  // namespace

void zgemm_allow_3m(int on) { g_allow_3m = on; }

void launch_zgemm_cb(int ta, int tb, int M, int N, int K, cplx alpha, const cplx* A, size_t lda, const cplx* B,
                     size_t ldb, cplx beta, cplx* C, size_t ldc, int lower, int batch, size_t sA, size_t sB,
                     size_t sC, int cb0, int cbs, int ncb, cudaStream_t st) {
  if (M <= 0 || N <= 0 || batch <= 0) return;
  const SplitK sk{};
  if (ta == 0 && tb == 0) launch_t<0, 0>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, lower, batch, sA, sB, sC, cb0, cbs, ncb, sk, st);
  else if (ta == 0 && tb == 1) launch_t<0, 1>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, lower, batch, sA, sB, sC, cb0, cbs, ncb, sk, st);
  else if (ta == 1 && tb == 0) launch_t<1, 0>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, lower, batch, sA, sB, sC, cb0, cbs, ncb, sk, st);
  else launch_t<1, 1>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, lower, batch, sA, sB, sC, cb0, cbs, ncb, sk, st);
}

void launch_zgemm_splitk(int ta, int tb, int M, int N, int K, const cplx* A, size_t lda, const cplx* B, size_t ldb,
                         cplx* Cparts, size_t ldc, size_t sC, int nseg, const SplitK& sk, cudaStream_t st) {
  if (M <= 0 || N <= 0 || nseg <= 0 || sk.chunks <= 0) return;
  const int batch = nseg * sk.chunks;
  const cplx one = cmake(1, 0), zero = cmake(0, 0);
  if (ta == 0 && tb == 0) launch_t<0, 0>(M, N, K, one, A, lda, B, ldb, zero, Cparts, ldc, 0, batch, 0, 0, sC, 0, 1, -1, sk, st);
  else if (ta == 0 && tb == 1) launch_t<0, 1>(M, N, K, one, A, lda, B, ldb, zero, Cparts, ldc, 0, batch, 0, 0, sC, 0, 1, -1, sk, st);
  else if (ta == 1 && tb == 0) launch_t<1, 0>(M, N, K, one, A, lda, B, ldb, zero, Cparts, ldc, 0, batch, 0, 0, sC, 0, 1, -1, sk, st);
  else launch_t<1, 1>(M, N, K, one, A, lda, B, ldb, zero, Cparts, ldc, 0, batch, 0, 0, sC, 0, 1, -1, sk, st);
}

void launch_zgemm(int ta, int tb, int M, int N, int K, cplx alpha, const cplx* A, size_t lda, const cplx* B,
                  size_t ldb, cplx beta, cplx* C, size_t ldc, int lower, int batch, size_t sA, size_t sB,
                  size_t sC, cudaStream_t st) {
  launch_zgemm_cb(ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, lower, batch, sA, sB, sC, 0, 1, -1, st);
}

}  // namespace zq
