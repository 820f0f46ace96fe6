// K2/K3: panel column kernels of the quaternion Householder tridiagonalisation.
//
// Replaces the k-loop of ts::impl::panel_update (reference blocked.cc:70-414: lazy column
// update :73-164, zlarfg :176/:345, zlartg :237, growth of T/R/S/W/Y/Z :171-413) with ONE
// quaternion reflector per column.  Per column k = j0 + i THREE launches:
//   col_update      x = M[:,k] - V W[k,:]^* - W V[k,:]^*   (and finishes w of column k-1);
//                   the LAST CTA to finish sums the norm partials and forms the reflector scalars
//                   (zlarfg analogue, tau real): record x[xrec..xrec+3) = (d,e,tau), alpha, u1^{-1}
//   matvec (K1)     v = x u1^{-1} is formed on the fly by every consumer (never a separate pass);
//                   partial sums of M v  + partial W^H v, V^H v
//   reduce_correct  p = tau (M v - V (W^H v) - W (V^H v)),  partial Re(v^H p); stores v into the panel
//                   and into the zeroed part of column k of A (reflector storage)
// All cross-CTA reductions go through small partial buffers summed in a fixed order, so the
// result is bit-reproducible run to run.
//
// Multi-GPU (1-D block-cyclic columns, one process per GPU).  ONE exchange per column and one per panel, both as
// peer-memory stores issued by these kernels (CUDA IPC + NVLink, PeerX in kernels.h):
//   * per panel: the owner of the panel's 64 columns pushes them (rows j0.., as they stand after the last trailing
//     update) into every rank's landing buffer (push_panel); from there EVERY rank forms x, the reflector scalars and v
//     of every column of the panel itself -- replicated, bit-identical work that was on the critical path of the
//     owner anyway -- so no per-column broadcast of the reflector exists;
//   * per column: reduce_correct pushes its 32 rows of the local partial M v into slot [rank] of every rank, raises a
//     per-row-block flag there, waits for the flags of the same row block from all ranks and sums the slots in rank
//     order (bit-identical on every rank): no grid-wide barrier, no separate kernel, no NCCL call.
//
// These kernels are latency-bound (n of each per solve): a CTA owns only PANEL_ROWS = 32 rows
// (lane = row, coalesced column-major access) and its 8 warps split the inner loops over the
// panel columns / partial buffers, so even a 4096-row column spreads over 128 CTAs.
#include "kernels.h"

namespace zq {

namespace {

constexpr int NT = 256, NW = 8, PR = PANEL_ROWS;

ZQ_D cplx* pan_ptr(const PanelWs& w, int which, int t) { return w.pan + ((size_t)(which * w.nb + t)) * w.n; }

// deterministic block-parallel sum of a small global array (all threads get the result)
ZQ_D double sum_parts(const double* __restrict__ p, int np, double* sm) {
  double s = 0.0;
  for (int j = threadIdx.x; j < np; j += NT) s += p[j];
  double v1[1] = {s};
  block_sum<1>(v1, sm);
  return v1[0];
}
// same, reading at L2 (partials written by other CTAs of the SAME grid)
ZQ_D double sum_parts_cg(const double* p, int np, double* sm) {
  double s = 0.0;
  for (int j = threadIdx.x; j < np; j += NT) s += __ldcg(p + j);
  double v1[1] = {s};
  block_sum<1>(v1, sm);
  return v1[0];
}

// cross-warp sum of one quat per (warp, lane): result valid in warp 0
ZQ_D quat warps_sum(quat v, quat (*sm)[PR], int warp, int lane) {
  sm[warp][lane] = v;
  __syncthreads();
  quat s = sm[0][lane];
  if (warp == 0) {
#pragma unroll
    for (int wv = 1; wv < NW; ++wv) s = qadd(s, sm[wv][lane]);
  }
  return s;
}

// ---- peer-exchange primitives (multi-GPU) ----------------------------------------------------
ZQ_D unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
ZQ_D void st_relaxed_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// wait (bounded) until *flag >= seq
ZQ_D void px_wait(const unsigned long long* flag, unsigned long long seq, int* info) {
  if (*(volatile int*)info & 8) return;                                  // exchange already failed: do not wait again
  const long long t0 = clock64();
  while (ld_acquire_sys(flag) < seq) {
    if (clock64() - t0 > 8000000000LL) { atomicOr(info, 8); break; }   // ~4 s: report instead of hanging the GPU
  }
}

// reflector scalars from ||x[k+2:]||^2 and x[k+1]:  alpha = -phase(x1) ||x||,  u1 = x1 - alpha,  tau = 2|u1|^2/||u||^2
struct Refl { quat alpha, inv; double tau, nx; };
ZQ_D Refl make_reflector(double rest2, quat x1) {
  Refl f;
  f.alpha = qzero(); f.inv = qzero(); f.tau = 0.0; f.nx = 0.0;
  const double x1n2 = qnorm2(x1);
  const double nx2 = rest2 + x1n2;
  if (nx2 > 0.0) {
    f.nx = sqrt(nx2);
    const double x1n = sqrt(x1n2);
    const quat ph = (x1n > 0.0) ? qscale(x1, 1.0 / x1n) : qmake(cmake(1, 0), cmake(0, 0));
    f.alpha = qscale(ph, -f.nx);
    const double u1n = x1n + f.nx;                 // u1 = x1 - alpha = phase * (|x1| + ||x||)
    const double u1n2 = u1n * u1n;
    f.tau = 2.0 * u1n2 / (rest2 + u1n2);
    f.inv = qscale(qconj(ph), 1.0 / u1n);          // u1^{-1}
  }
  return f;
}

// ---------------------------------------------------------------------------------------------
// col_update: rows r in [k, n).  Column k of the matrix as of the start of the panel is read from A or, on several
// GPUs, from the panel landing buffer w.apanel (only the owner's copy of A is current).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT) k_col_update(PanelWs w, int k, int j0, int ng_parts) {
  pdl_enter();
  const int n = w.n, i = k - j0;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r = k + blockIdx.x * PR + lane;
  __shared__ quat qW[MAX_NB_PANEL], qV[MAX_NB_PANEL];
  __shared__ quat red[NW][PR];
  __shared__ double s_red[32];
  __shared__ int s_last;
  const bool act = r < n;

  quat wr = qzero();       // W(r, i-1) for this thread's row
  if (i > 0) {
    // finish w of the previous column: w = p - 1/2 tau (v^H p) v   (zlatrd-style; tau real)
    const double g = sum_parts(w.g_part, ng_parts, s_red);
    const double coef = 0.5 * w.tau[k - 1] * g;
    const cplx* va = pan_ptr(w, 0, i - 1);
    const cplx* vb = pan_ptr(w, 1, i - 1);
    if (act) {
      quat pr = w.p[r];
      quat vr = qmake(va[r], vb[r]);
      wr = qmake(csub(pr.a, cscale(vr.a, coef)), csub(pr.b, cscale(vr.b, coef)));
      if (warp == 0) {
        pan_ptr(w, 2, i - 1)[r] = wr.a;
        pan_ptr(w, 3, i - 1)[r] = wr.b;
      }
    }
    // row-k coefficients: qconj(W(k,t)), qconj(V(k,t)), t < i
    for (int t = threadIdx.x; t < i; t += NT) {
      quat wk, vk;
      if (t == i - 1) {
        quat pk = w.p[k];
        quat vkk = qmake(va[k], vb[k]);
        wk = qmake(csub(pk.a, cscale(vkk.a, coef)), csub(pk.b, cscale(vkk.b, coef)));
        vk = vkk;
      } else {
        wk = qmake(pan_ptr(w, 2, t)[k], pan_ptr(w, 3, t)[k]);
        vk = qmake(pan_ptr(w, 0, t)[k], pan_ptr(w, 1, t)[k]);
      }
      qW[t] = qconj(wk);
      qV[t] = qconj(vk);
    }
  }
  __syncthreads();

  quat part = qzero();     // - sum over this warp's panel columns
  quat acol = qzero();     // A(r, k): issued before the panel loop so its HBM latency hides behind it
  if (warp == 0 && act) {
    if (w.apanel) {        // landing buffer [part][panel column][row], written by the panel's owner over NVLink: read at L2
      const double2* pa = reinterpret_cast<const double2*>(w.apanel + (size_t)i * w.apanel_ld + r);
      const double2* pb = reinterpret_cast<const double2*>(w.apanel + ((size_t)MAX_NB_PANEL + i) * w.apanel_ld + r);
      acol = qmake(__ldcg(pa), __ldcg(pb));
    } else {
      acol = qmake(w.A[(size_t)r + (size_t)k * w.lda], w.A[(size_t)(n + r) + (size_t)k * w.lda]);
    }
  }
  if (act) {
#pragma unroll 4
    for (int t = warp; t < i; t += NW) {
      quat vrt = qmake(pan_ptr(w, 0, t)[r], pan_ptr(w, 1, t)[r]);
      quat wrt = (t == i - 1) ? wr : qmake(pan_ptr(w, 2, t)[r], pan_ptr(w, 3, t)[r]);
      qfms(part, vrt, qW[t]);
      qfms(part, wrt, qV[t]);
    }
  }
  quat col = warps_sum(part, red, warp, lane);
  double nr = 0.0;
  if (warp == 0 && act) {
    col = qadd(col, acol);
    if (r == k) {
      w.d[k] = col.a.x;
    } else {
      w.x[r] = col;
      if (r >= k + 2) nr = qnorm2(col);
    }
  }
  if (warp == 0) {
    nr = warp_sum(nr);
    if (lane == 0) w.nrm_part[blockIdx.x] = nr;
  }
  if (k + 1 >= n) return;                         // last diagonal entry only: no reflector

  // ---- the last CTA to arrive closes the column: reflector scalars -> record ----
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned prev = atomicAdd(w.counter, 1u);
    s_last = (prev == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return;
  if (threadIdx.x == 0) *(volatile unsigned int*)w.counter = 0u;
  __threadfence();
  const double rest2 = sum_parts_cg(w.nrm_part, gridDim.x, s_red);
  if (threadIdx.x == 0) {
    const double2* xp = reinterpret_cast<const double2*>(w.x + k + 1);
    const quat x1 = qmake(__ldcg(xp), __ldcg(xp + 1));
    const Refl f = make_reflector(rest2, x1);
    w.x[w.xrec] = qmake(cmake(__ldcg(w.d + k), f.nx), cmake(f.tau, 0.0));
    w.x[w.xrec + 1] = f.alpha;
    w.x[w.xrec + 2] = f.inv;
  }
}

// ---------------------------------------------------------------------------------------------
// push_panel (multi-GPU, once per panel, owner only): rows [j0, n) of the panel's kb columns of D and E -> the landing
// buffer of EVERY rank (own included), then the sequence number -> the peers' panel flag.
// wait_flag (the other ranks): one thread waits until the flag has arrived.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT) k_push_panel(PanelWs w, PeerX px, int j0, int kb, unsigned long long seq) {
  const int n = w.n, t = blockIdx.y;
  const int r = j0 + blockIdx.x * NT + threadIdx.x;
  __shared__ int s_last;
  if (r < n) {
    const cplx d = w.A[(size_t)r + (size_t)(j0 + t) * w.lda];
    const cplx e = w.A[(size_t)(n + r) + (size_t)(j0 + t) * w.lda];
    const size_t ia = (size_t)t * px.nmax + r, ib = ((size_t)MAX_NB_PANEL + t) * px.nmax + r;
    for (int g = 0; g < px.world; ++g) {
      px.apanel[g][ia] = d;
      px.apanel[g][ib] = e;
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned prev = atomicAdd(w.counter, 1u);
    s_last = (prev == gridDim.x * gridDim.y - 1);
  }
  __syncthreads();
  if (!s_last) return;
  if (threadIdx.x == 0) {
    *(volatile unsigned int*)w.counter = 0u;
    __threadfence_system();
    for (int g = 0; g < px.world; ++g)
      if (g != px.rank) st_relaxed_sys(px.flags[g] + 0, seq);
  }
}

__global__ void k_wait_flag(PeerX px, int idx, unsigned long long seq) {
  if (threadIdx.x == 0) px_wait(px.flags[px.rank] + idx, seq, px.info);
}

// local variant for the NCCL transport: the owner packs the panel's columns into its OWN buffer (then ncclBroadcast)
__global__ void __launch_bounds__(NT) k_pack_panel(PanelWs w, cplx* buf, size_t ld, int j0) {
  const int n = w.n, t = blockIdx.y;
  const int r = j0 + blockIdx.x * NT + threadIdx.x;
  if (r >= n) return;
  buf[(size_t)t * ld + r] = w.A[(size_t)r + (size_t)(j0 + t) * w.lda];
  buf[((size_t)MAX_NB_PANEL + t) * ld + r] = w.A[(size_t)(n + r) + (size_t)(j0 + t) * w.lda];
}

// v[r] of the current column from x and the scalar record (read at L2: they may have been written by a peer GPU
// or by another CTA's last-block epilogue)
ZQ_D quat ldcg_quat(const quat* p) {
  const double2* q = reinterpret_cast<const double2*>(p);
  return qmake(__ldcg(q), __ldcg(q + 1));
}

// ---------------------------------------------------------------------------------------------
// reduce_correct: rows r in [k+1, n)
// ---------------------------------------------------------------------------------------------
// MODE 0: single GPU (sum all K1 partials, correct, scale).  MODE 1: NCCL transport, local part only:
// w.p[r] = sum of THIS rank's partials.  MODE 2: NCCL transport, after the all-reduce of w.p: correct, scale.
// MODE 3: peer exchange, everything in one kernel: local part -> slot [rank] of every rank's staging buffer +
// per-row-block flag; wait for the same row block of all ranks; sum in rank order; correct; scale.
// All modes except 1 also unpack the column: d, e, tau, alpha from the scalar record and v = x u1^{-1} into the
// panel and into the reflector storage of A.
template <int MODE>
__global__ void __launch_bounds__(NT) k_reduce_correct(PanelWs w, PeerX px, int k, int j0, int nch, int tpb, unsigned long long seq) {
  pdl_enter();
  constexpr bool PART = (MODE == 1), FIN = (MODE == 2);
  const int n = w.n, i = k - j0, s = k + 1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r = s + blockIdx.x * PR + lane;
  __shared__ quat gW[MAX_NB_PANEL], gV[MAX_NB_PANEL];
  __shared__ quat red[NW][PR];
  if (!PART) {
    for (int t = threadIdx.x; t < 2 * i; t += NT) {
      const int tt = t >> 1;
      const quat* src = (t & 1) ? w.dotV : w.dotW;
      quat acc = qzero();
#pragma unroll 8
      for (int c = 0; c < nch; ++c) acc = qadd(acc, src[(size_t)c * w.nb + tt]);
      if (t & 1) gV[tt] = acc; else gW[tt] = acc;
    }
    __syncthreads();
    if (blockIdx.x == 0) {
      for (int t = threadIdx.x; t < i; t += NT) w.G[(size_t)k * w.nb + t] = gV[t];
    }
  }
  quat part = qzero();
  if (r < n && !FIN) {
    const int J0 = s / MV_TC, Jlast = (n - 1) / MV_TC;
    const int I0 = s / MV_TR, I1 = (n - 1) / MV_TR;
    const int Jhi = min(2 * (r / MV_TR) + 1, Jlast);
    const int Ilo = max(I0, (r / MV_TC) / 2);
    // direct sums: one slot per RUN of tpb owned column blocks (matvec.cu); owned blocks J = rank + q*world
    const int G = w.world;
    const int jfirst = J0 + ((w.rank - J0 % G) + G) % G;
    if (jfirst <= Jhi) {
      const int Rlo = (jfirst / G) / tpb, Rhi = ((Jhi - w.rank) / G) / tpb;
#pragma unroll 4
      for (int R = Rlo + warp; R <= Rhi; R += NW) part = qadd(part, w.pd[(size_t)R * n + r]);
    }
    if ((r / MV_TC) % w.world == w.rank) {   // transposed sums exist only on the owner of r's column block
#pragma unroll 4
      for (int I = Ilo + warp; I <= I1; I += NW) part = qadd(part, w.pt[(size_t)I * n + r]);
    }
  }
  if (MODE == 1) {
    const quat y = warps_sum(part, red, warp, lane);
    if (warp == 0 && r < n) w.p[r] = y;
    return;
  }
  if (MODE == 3) {
    // ---- exchange of this row block: push, flag, wait, ordered sum ----
    const quat yl = warps_sum(part, red, warp, lane);
    const int par = k & 1;
    if (warp == 0) {
      if (r < n) {
        const size_t at = ((size_t)par * px.world + px.rank) * px.nmax + r;
        for (int g = 0; g < px.world; ++g) px.ypart[g][at] = yl;        // own slot on every rank (NVLink stores)
      }
      __threadfence_system();
      __syncwarp();
      const size_t fl = 8 + ((size_t)par * px.rbmax + blockIdx.x) * PX_MAXW;
      if (lane < px.world) st_relaxed_sys(px.flags[lane] + fl + px.rank, seq);
      if (lane < px.world) px_wait(px.flags[px.rank] + fl + lane, seq, px.info);
      __syncwarp();
      part = qzero();
      if (r < n) {
        const quat* yp = px.ypart[px.rank] + (size_t)par * px.world * px.nmax;
        for (int g = 0; g < px.world; ++g) part = qadd(part, ldcg_quat(yp + (size_t)g * px.nmax + r));
      }
    } else {
      part = qzero();
    }
    __syncthreads();                              // `red` is reused below
  }
  if (FIN) {
    if (warp == 0 && r < n) part = w.p[r];        // all-reduced M v (NCCL)
  }
  if (r < n) {
#pragma unroll 4
    for (int t = warp; t < i; t += NW) {
      quat vrt = qmake(pan_ptr(w, 0, t)[r], pan_ptr(w, 1, t)[r]);
      quat wrt = qmake(pan_ptr(w, 2, t)[r], pan_ptr(w, 3, t)[r]);
      qfms(part, vrt, gW[t]);
      qfms(part, wrt, gV[t]);
    }
  }
  quat y = warps_sum(part, red, warp, lane);
  if (warp == 0) {
    const quat sc = ldcg_quat(w.x + w.xrec);
    double g = 0.0;
    if (r < n) {
      const double tau = sc.b.x;
      y = qscale(y, tau);
      w.p[r] = y;
      // v of this column: stored where the next kernels (and the back-transformation) read it
      quat f;
      if (r == s) {
        f = qmake(cmake(1, 0), cmake(0, 0));
      } else {
        f = qmul(ldcg_quat(w.x + r), ldcg_quat(w.x + w.xrec + 2));     // right scaling keeps H = I - tau v v^H Hermitian
        w.A[(size_t)r + (size_t)k * w.lda] = f.a;                       // reflector tail lives where zeros were created
        w.A[(size_t)(n + r) + (size_t)k * w.lda] = f.b;
      }
      pan_ptr(w, 0, i)[r] = f.a;
      pan_ptr(w, 1, i)[r] = f.b;
      g = f.a.x * y.a.x + f.a.y * y.a.y + f.b.x * y.b.x + f.b.y * y.b.y;   // Re(f^H y)
    }
    g = warp_sum(g);
    if (lane == 0) w.g_part[blockIdx.x] = g;
    if (blockIdx.x == 0 && lane == 0) {
      w.d[k] = sc.a.x;
      w.e[k] = sc.a.y;
      w.tau[k] = sc.b.x;
      w.alpha[k] = ldcg_quat(w.x + w.xrec + 1);
    }
  }
}

// finish w of the LAST column of a panel (col_update does it for the others)
__global__ void __launch_bounds__(NT) k_finish_w(PanelWs w, int k, int j0, int ng_parts) {
  pdl_enter();
  const int n = w.n, i = k - j0;
  const int r = k + 1 + blockIdx.x * NT + threadIdx.x;
  __shared__ double s_red[32];
  const double g = sum_parts(w.g_part, ng_parts, s_red);
  const double coef = 0.5 * w.tau[k] * g;
  if (r < n) {
    quat pr = w.p[r];
    quat vr = qmake(pan_ptr(w, 0, i)[r], pan_ptr(w, 1, i)[r]);
    pan_ptr(w, 2, i)[r] = csub(pr.a, cscale(vr.a, coef));
    pan_ptr(w, 3, i)[r] = csub(pr.b, cscale(vr.b, coef));
  }
}

// K4 operand staging: L = [[Va, conj Vb, Wa, conj Wb], [Vb, -conj Va, Wb, -conj Wa]],
//                     R = [Wa, conj Wb, Va, conj Vb]   (rows r0..n-1)
__global__ void __launch_bounds__(256) k_build_LR(PanelWs w, int r0, int kb, cplx* L, cplx* R) {
  const int n = w.n, m = n - r0;
  const int t = blockIdx.y;
  const int rr = blockIdx.x * 256 + threadIdx.x;
  if (rr >= m) return;
  const int r = r0 + rr;
  const cplx va = pan_ptr(w, 0, t)[r], vb = pan_ptr(w, 1, t)[r];
  const cplx wa = pan_ptr(w, 2, t)[r], wb = pan_ptr(w, 3, t)[r];
  const size_t ldl = 2 * (size_t)m, ldr = m;
  L[rr + (size_t)(0 * kb + t) * ldl] = va;
  L[rr + (size_t)(1 * kb + t) * ldl] = cconj(vb);
  L[rr + (size_t)(2 * kb + t) * ldl] = wa;
  L[rr + (size_t)(3 * kb + t) * ldl] = cconj(wb);
  L[m + rr + (size_t)(0 * kb + t) * ldl] = vb;
  L[m + rr + (size_t)(1 * kb + t) * ldl] = cneg(cconj(va));
  L[m + rr + (size_t)(2 * kb + t) * ldl] = wb;
  L[m + rr + (size_t)(3 * kb + t) * ldl] = cneg(cconj(wa));
  R[rr + (size_t)(0 * kb + t) * ldr] = wa;
  R[rr + (size_t)(1 * kb + t) * ldr] = cconj(wb);
  R[rr + (size_t)(2 * kb + t) * ldr] = va;
  R[rr + (size_t)(3 * kb + t) * ldr] = cconj(vb);
}

// quaternion form of the K4 operands (qgemm.cu): Aq = [V W], Sq = [W V], a-part rows [0, m), b-part rows [m, 2m)
__global__ void __launch_bounds__(256) k_build_VW(PanelWs w, int r0, int kb, cplx* Aq, cplx* Sq) {
  const int n = w.n, m = n - r0;
  const int t = blockIdx.y;
  const int rr = blockIdx.x * 256 + threadIdx.x;
  if (rr >= m) return;
  const int r = r0 + rr;
  const cplx va = pan_ptr(w, 0, t)[r], vb = pan_ptr(w, 1, t)[r];
  const cplx wa = pan_ptr(w, 2, t)[r], wb = pan_ptr(w, 3, t)[r];
  const size_t ld = 2 * (size_t)m;
  Aq[rr + (size_t)t * ld] = va;          Aq[m + rr + (size_t)t * ld] = vb;
  Aq[rr + (size_t)(kb + t) * ld] = wa;   Aq[m + rr + (size_t)(kb + t) * ld] = wb;
  Sq[rr + (size_t)t * ld] = wa;          Sq[m + rr + (size_t)t * ld] = wb;
  Sq[rr + (size_t)(kb + t) * ld] = va;   Sq[m + rr + (size_t)(kb + t) * ld] = vb;
}

inline int cdiv(int a, int b) { return (a + b - 1) / b; }

}  // namespace

void launch_col_update(const PanelWs& w, int k, int j0, cudaStream_t st) {
  const int rows = w.n - k;                    // rows k..n-1
  const int ng = (k > j0) ? cdiv(w.n - k, PR) : 0;   // g_part written by reduce_correct of column k-1 (rows k..n-1)
  launch_chain(k_col_update, dim3(cdiv(rows, PR)), dim3(NT), st, w, k, j0, ng);
}

void launch_push_panel(const PanelWs& w, const PeerX& px, int j0, int kb, unsigned long long seq, cudaStream_t st) {
  k_push_panel<<<dim3(cdiv(w.n - j0, NT), kb), NT, 0, st>>>(w, px, j0, kb, seq);
}

void launch_wait_panel(const PeerX& px, unsigned long long seq, cudaStream_t st) { k_wait_flag<<<1, 32, 0, st>>>(px, 0, seq); }

void launch_pack_panel(const PanelWs& w, cplx* buf, size_t ld, int j0, int kb, cudaStream_t st) {
  k_pack_panel<<<dim3(cdiv(w.n - j0, NT), kb), NT, 0, st>>>(w, buf, ld, j0);
}

void launch_reduce_correct(const PanelWs& w, int k, int j0, cudaStream_t st) {
  const int rows = w.n - k - 1;
  const int nch = cdiv(rows, dot_chunk_rows(rows));
  launch_chain(k_reduce_correct<0>, dim3(cdiv(rows, PR)), dim3(NT), st, w, PeerX{}, k, j0, nch, k1_tpb(rows, w.world), 0ull);
}

void launch_reduce_partial(const PanelWs& w, int k, cudaStream_t st) {
  const int rows = w.n - k - 1;
  launch_chain(k_reduce_correct<1>, dim3(cdiv(rows, PR)), dim3(NT), st, w, PeerX{}, k, k, 0, k1_tpb(rows, w.world), 0ull);
}

void launch_correct(const PanelWs& w, int k, int j0, cudaStream_t st) {
  const int rows = w.n - k - 1;
  const int nch = cdiv(rows, dot_chunk_rows(rows));
  launch_chain(k_reduce_correct<2>, dim3(cdiv(rows, PR)), dim3(NT), st, w, PeerX{}, k, j0, nch, 1, 0ull);
}

void launch_reduce_correct_px(const PanelWs& w, const PeerX& px, int k, int j0, unsigned long long seq, cudaStream_t st) {
  const int rows = w.n - k - 1;
  launch_chain(k_reduce_correct<3>, dim3(cdiv(rows, PR)), dim3(NT), st, w, px, k, j0, cdiv(rows, dot_chunk_rows(rows)), k1_tpb(rows, w.world), seq);
}

void launch_finish_w(const PanelWs& w, int k, int j0, cudaStream_t st) {
  const int rows = w.n - k - 1;
  launch_chain(k_finish_w, dim3(cdiv(rows, NT)), dim3(NT), st, w, k, j0, cdiv(rows, PR));
}

void launch_build_VW(const PanelWs& w, int r0, int kb, cplx* Aq, cplx* Sq, cudaStream_t st) {
  const int m = w.n - r0;
  dim3 g(cdiv(m, 256), kb);
  k_build_VW<<<g, 256, 0, st>>>(w, r0, kb, Aq, Sq);
}

void launch_build_LR(const PanelWs& w, int r0, int kb, cplx* L, cplx* R, cudaStream_t st) {
  const int m = w.n - r0;
  dim3 g(cdiv(m, 256), kb);
  k_build_LR<<<g, 256, 0, st>>>(w, r0, kb, L, R);
}

}  // namespace zq
