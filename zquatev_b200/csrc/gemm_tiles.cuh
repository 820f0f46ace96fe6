// Building blocks shared by the FP64 tensor-path GEMM kernels (zgemm.cu: complex 4-product / 3M; qgemm.cu: quaternion
// 8-product): cp.async staging, padded shared-memory operand tiles, the per-thread copy schedule and the DMMA wrapper.
#pragma once
#include "kernels.h"

namespace zq {

ZQ_D void cp_async16(void* smem, const void* gmem, bool pred) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  const int sz = pred ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(sz));
}
ZQ_D void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
ZQ_D void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

ZQ_D void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// Operand tile in shared memory.  KCONT = false: element (t, k) at k*LD + t, LD = BT + 2
// (global storage contiguous along the tile dimension t);  KCONT = true: element (t, k) at
// t*LD + k, LD = BK + 4 (global storage contiguous along k).  Both paddings make the 16-byte
// fragment loads of a quarter warp hit 8 distinct 16-byte bank groups.
template <int BT, bool KCONT, int BK>
struct OpTile {
  static constexpr int LD = KCONT ? (BK + 4) : (BT + 2);
  static constexpr int ELEMS = KCONT ? BT * LD : BK * LD;
  // global element (t, k): KCONT ? G[k + t*ldg] : G[t + k*ldg]
  template <int NT>
  static ZQ_D void load(cplx* sm, const cplx* __restrict__ G, size_t ldg, int t0, int k0, int Tmax, int Kmax, int tid) {
#pragma unroll
    for (int e = tid; e < BT * BK; e += NT) {
      int t, k;
      if (KCONT) { k = e % BK; t = e / BK; } else { t = e % BT; k = e / BT; }
      const bool ok = (t0 + t < Tmax) && (k0 + k < Kmax);
      const cplx* src = ok ? (KCONT ? G + (size_t)(k0 + k) + (size_t)(t0 + t) * ldg
                                    : G + (size_t)(t0 + t) + (size_t)(k0 + k) * ldg) : G;
      cp_async16(sm + (KCONT ? t * LD + k : k * LD + t), src, ok);
    }
  }
  static ZQ_D cplx frag(const cplx* sm, int t, int k) { return sm[KCONT ? t * LD + k : k * LD + t]; }
};

// Per-thread cp.async schedule of one operand tile.  Of a thread's (t, k) elements one index is the same for
// every element and the other advances by a constant, so everything except the k-tail test is computed once
// before the main loop: a stage costs one predicate, one 64-bit add and one LDGSTS per 16 bytes (the generic
// index arithmetic of OpTile::load was ~25 integer instructions per copy, issued in front of every DMMA block).
template <int BT, bool KCONT, int BK, int NT>
struct TileLoader {
  using Tile = OpTile<BT, KCONT, BK>;
  static constexpr int PER = (BT * BK) / NT;                  // 16-byte copies per thread per stage
  static constexpr int ISTEP = KCONT ? NT / BK : NT / BT;     // distance of consecutive copies in the varying index
  static_assert((BT * BK) % NT == 0 && (KCONT ? NT % BK == 0 : NT % BT == 0), "tile / thread count mismatch");
  const cplx* p0;      // source of copy 0 of the next stage
  const cplx* safe;    // any valid address (zero-fill copies read nothing)
  size_t istride;      // elements between consecutive copies
  size_t kstep;        // elements between consecutive stages
  int s0;              // shared-memory element offset of copy 0
  int kfix;            // k of copy 0 inside the stage
  unsigned okmask;     // bit i: copy i is inside the tile dimension
  ZQ_D void init(const cplx* G, size_t ldg, int t0, int Tmax, int tid) {
    safe = G;
    istride = (size_t)ISTEP * ldg;
    if (KCONT) {
      const int k = tid % BK, tb = tid / BK;
      p0 = G + (size_t)k + (size_t)(t0 + tb) * ldg;
      kstep = BK;
      s0 = tb * Tile::LD + k;
      kfix = k;
      okmask = 0;
#pragma unroll
      for (int i = 0; i < PER; ++i) okmask |= (t0 + tb + i * ISTEP < Tmax) ? (1u << i) : 0u;
    } else {
      const int t = tid % BT, kb = tid / BT;
      p0 = G + (size_t)(t0 + t) + (size_t)kb * ldg;
      kstep = (size_t)BK * ldg;
      s0 = kb * Tile::LD + t;
      kfix = kb;
      okmask = (t0 + t < Tmax) ? 0xffffffffu : 0u;
    }
  }
  // stage whose first k is k0 (stages must be issued in order: p0 advances)
  ZQ_D void issue(cplx* sm, int k0, int K) {
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int k = KCONT ? kfix : kfix + i * ISTEP;
      const bool ok = ((okmask >> i) & 1u) && (k0 + k < K);
      cp_async16(sm + s0 + i * (ISTEP * Tile::LD), ok ? p0 + (size_t)i * istride : safe, ok);
    }
    p0 += kstep;
  }
};

// the eight left / right combinations of the four components (CONJ: quaternion conjugate of the stored element)
template <bool CONJ>
ZQ_D void combos_a(cplx a, cplx b, double (&al)[8]) {
  const double q0 = a.x, q1 = CONJ ? -a.y : a.y, q2 = CONJ ? -b.x : b.x, q3 = CONJ ? b.y : -b.y;
  al[0] = q3 + q1; al[1] = q0 - q2; al[2] = q0 + q2; al[3] = q3 - q1;
  al[4] = q3 - q2; al[5] = q1 + q0; al[6] = q0 - q1; al[7] = q3 + q2;
}
template <bool CONJ>
ZQ_D void combos_b(cplx a, cplx b, double (&be)[8]) {
  const double q0 = a.x, q1 = CONJ ? -a.y : a.y, q2 = CONJ ? -b.x : b.x, q3 = CONJ ? b.y : -b.y;
  be[0] = q1 + q2; be[1] = q0 + q3; be[2] = q0 - q3; be[3] = q1 - q2;
  be[4] = q2 - q3; be[5] = q1 + q0; be[6] = q2 + q3; be[7] = q0 - q1;
}

}  // namespace zq
