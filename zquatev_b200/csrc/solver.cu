// Host orchestrator + C ABI (include/zquatev_b200.h).
//
// Mirrors the driver of the reference, ts::zquatev (zquatev.cc:42-100):
//   reference                                   here
//   repack D0/D1, Q0=I, Q1=0   (:48-61)         none: kernels index (A;B) in place; Q never formed
//   panel_update / unblocked   (:68-72)         tridiagonalise(): col_update, reflector, K1, reduce_correct, K4
//   band pack + zhbev          (:76-84)         phase chain + dc_solve() (K8) or bisection (K9)
//   U = Q0 Z, V = Q1 Z         (:87-90)         backtransform(): compact-WY GEMMs on X0 = diag(s) Z (K6)
//   symmetry fill              (:93-98)         swap_pairing (K10)
// Everything is enqueued on one stream; the host never waits inside a solve except for the
// final status word.
#include <complex>
#include <cstdio>
#include <cstring>
#include <atomic>
#include <mutex>
#include <thread>
#include <vector>
#include <dlfcn.h>
#include <nccl.h>
#include <nvtx3/nvToolsExt.h>
#include "kernels.h"
#include "../../include/zquatev_b200.h"

int zq_cuda_fail(cudaError_t e, const char* file, int line) {
  fprintf(stderr, "[zquatev_b200] CUDA error %d (%s) at %s:%d\n", (int)e, cudaGetErrorString(e), file, line);
  return -(1000 + (int)e);
}

namespace zq {

constexpr int DEFAULT_NB = 64;
constexpr int MAX_NB = 64;
constexpr int YP_PARTS = 8;        // split-K workspace of the back-transformation: 8 partial Y at ncols = n
constexpr int YP_MAX_CHUNKS = 8;   // at most 8 K-chunks per (a-rows / b-rows) segment

struct Plan {
  int n = 0, nb = 0, device = -1;
  char* slab = nullptr;
  PanelWs pw{};
  cplx *L = nullptr, *R = nullptr, *P = nullptr, *T = nullptr, *Y = nullptr, *TY = nullptr, *YP = nullptr;
  cplx *T12 = nullptr, *S12 = nullptr, *ST = nullptr;   // two-panel back-transformation (ZQ_BT_PAIR)
  size_t yp_elems = 0;       // capacity of YP (split-K partial products of Y = Phi(V)^H X)
  quat* s = nullptr;
  double* bis = nullptr;
  double* dcv = nullptr;     // n + 8*PX_MAXW doubles: all-gather staging of the sharded bisection
  double* scl = nullptr;     // input scaling: (sigma, 1/sigma) + partial maxima
  int* info_dev = nullptr;
  DcWs* dc = nullptr;
  cplx* Dfull = nullptr;     // host-pointer mode staging, 2n x 2n
  cplx* apanel = nullptr;    // multi-GPU over the NCCL transport: the current panel's columns [2][64][n] (allocated on first use)
  double *X8a = nullptr, *X8b = nullptr;   // pre-combined operands of the ZQ_Q8X GEMM variant (qgemm8x.cu; allocated on first use)
  double* eig_dev = nullptr;
  cudaEvent_t ev[6] = {};
  cudaEvent_t ev_gather = nullptr;   // multi-GPU: recorded before the NCCL gather of the eigenvector shards
  cudaStream_t cs = nullptr;         // host-pointer mode: D2H of finished eigenvector column blocks overlaps the next block
  cudaEvent_t ev_chunk[4] = {};      //   block c back-transformed (recorded on the solver's stream, awaited by cs)
  cudaEvent_t ev_copy = nullptr;     //   all downloads issued on cs are complete
  cudaStream_t gs = nullptr;         // collective host-pointer mode: NCCL gather of the ranks' finished column sub-blocks (so that a
  cudaEvent_t ev_gath[4] = {};       //   rank busy downloading never delays the others' sends); sub-block c gathered
  cudaEvent_t done = nullptr;        // recorded at the end of every solve: the next user of this workspace waits on it,
  bool done_valid = false;           // so an asynchronous solve on another stream cannot race with it
  double gather_ms = 0;
  std::vector<cudaEvent_t> k1ev;
  std::vector<cudaEvent_t> k4ev;   // profiling: event pair around every trailing-update GEMM
  double k4_ms = 0;
  double phase_ms[8] = {};
  long launches = 0;
  bool timing = true;        // false while a solve is being captured into a CUDA graph (no event records)
};

// ---- NCCL, resolved at run time (no link dependency: single-GPU users never load it) ----------
struct NcclApi {
  void* h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;
static ncclComm_t g_comm = nullptr;
static int g_rank = 0, g_world = 1;
// peer-memory exchange state (CUDA IPC); g_px.world == 0: not available -> NCCL per-column collectives
static PeerX g_px{};
static char* g_xb = nullptr;                       // this rank's exchange buffer
static void* g_peer_base[PX_MAXW] = {};
static unsigned long long g_seq_base = 0;          // sequence numbers never restart (flags are never reset)

static int nccl_load() {
  if (g_nccl.h) return 0;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) { fprintf(stderr, "[zquatev_b200] cannot load libnccl: %s\n", dlerror()); return -900; }
#define ZQ_SYM(field, name) *(void**)(&g_nccl.field) = dlsym(h, name); if (!g_nccl.field) { fprintf(stderr, "[zquatev_b200] missing %s\n", name); return -901; }
  ZQ_SYM(GetUniqueId, "ncclGetUniqueId") ZQ_SYM(CommInitRank, "ncclCommInitRank") ZQ_SYM(CommDestroy, "ncclCommDestroy")
  ZQ_SYM(Broadcast, "ncclBroadcast") ZQ_SYM(AllReduce, "ncclAllReduce") ZQ_SYM(GroupStart, "ncclGroupStart")
  ZQ_SYM(GroupEnd, "ncclGroupEnd") ZQ_SYM(AllGather, "ncclAllGather") ZQ_SYM(GetErrorString, "ncclGetErrorString")
  ZQ_SYM(Send, "ncclSend") ZQ_SYM(Recv, "ncclRecv")
#undef ZQ_SYM
  g_nccl.h = h;
  return 0;
}

static int nccl_fail(ncclResult_t r, int line) {
  fprintf(stderr, "[zquatev_b200] NCCL error %d (%s) at solver.cu:%d\n", (int)r, g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?", line);
  return -(900 + (int)r);
}
#define ZQ_NCCL_CHECK(expr) do { ncclResult_t _r = (expr); if (_r != ncclSuccess) return nccl_fail(_r, __LINE__); } while (0)

// A handle owns a cache of plans (device workspaces, one per (n, nb)) on ONE device and its own lock: solves through
// different handles do not serialise on each other (the reference allocates and frees its workspace inside every
// call, zquatev.cc:63-66 "TODO"; here a caller that alternates between sizes keeps all of them warm).
struct Handle {
  int device = -1;
  std::mutex mu;
  std::vector<Plan*> plans;          // most recently used first
  Plan* last = nullptr;              // plan of the last solve (phase timings)
  size_t max_plans = 4;
};
static std::mutex g_mu;              // default-handle table, batched lanes, communicator set-up
static std::mutex g_dist_mu;         // collective solves share the one communicator of the process
static Handle* g_default[64] = {};   // default handle per device: the reference-shaped entries use it
static bool g_profile = false;

static void plan_free(Plan* p) {
  if (!p) return;
  if (p->slab) cudaFree(p->slab);
  if (p->Dfull) cudaFree(p->Dfull);
  if (p->apanel) cudaFree(p->apanel);
  if (p->X8a) cudaFree(p->X8a);
  if (p->X8b) cudaFree(p->X8b);
  if (p->dc) dc_destroy(p->dc);
  for (auto& e : p->ev) if (e) cudaEventDestroy(e);
  if (p->ev_gather) cudaEventDestroy(p->ev_gather);
  if (p->done) cudaEventDestroy(p->done);
  for (auto& e : p->ev_chunk) if (e) cudaEventDestroy(e);
  if (p->ev_copy) cudaEventDestroy(p->ev_copy);
  if (p->cs) cudaStreamDestroy(p->cs);
  for (auto& e : p->ev_gath) if (e) cudaEventDestroy(e);
  if (p->gs) cudaStreamDestroy(p->gs);
  for (auto& e : p->k1ev) cudaEventDestroy(e);
  for (auto& e : p->k4ev) cudaEventDestroy(e);
  delete p;
}

static int cdiv(int a, int b) { return (a + b - 1) / b; }

// K4/K6 as quaternion GEMMs with eight real products (qgemm.cu) instead of stacked complex 3M GEMMs: where the 3M scheme
// is allowed (n >= 1024); ZQ_QGEMM=0 keeps the stacked complex form.
static bool use_qgemm(int n) {
  static const int on = [] { const char* e = getenv("ZQ_QGEMM"); return e ? atoi(e) : 1; }();
  return on && n >= 1024;
}

// NVTX ranges around the phases of a solve as the host ENQUEUES them (ZQ_NVTX=1; header-only NVTX 3: a no-op unless a
// profiler injects itself).  A timeline tool projects them onto the kernels of the solver's stream.
struct NvtxRange {
  bool on;
  explicit NvtxRange(const char* name) {
    static const bool enabled = [] { const char* e = getenv("ZQ_NVTX"); return e && atoi(e) != 0; }();
    on = enabled;
    if (on) nvtxRangePushA(name);
  }
  ~NvtxRange() { if (on) nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};

// ZQ_Q8X=1: the trailing update and the update half of the back-transformation on the pre-combined-operand GEMM (qgemm8x.cu).
// Its two operand workspaces (8 planes x 128 x n doubles each) are allocated on first use; false: use k_qgemm8.
static bool use_q8x(Plan* p) {
  if (!qgemm_x_enabled() || !use_qgemm(p->n)) return false;
  if (!p->X8a || !p->X8b) {
    if (!p->timing) return false;                   // the solve is being captured into a CUDA graph: no allocation now
    const size_t bytes = qgemm_x_operand_doubles(p->n, 2 * MAX_NB) * sizeof(double);
    if (!p->X8a && cudaMalloc(&p->X8a, bytes) != cudaSuccess) { cudaGetLastError(); p->X8a = nullptr; return false; }
    if (!p->X8b && cudaMalloc(&p->X8b, bytes) != cudaSuccess) { cudaGetLastError(); p->X8b = nullptr; return false; }
  }
  return true;
}

bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("ZQ_PDL"); return e ? atoi(e) != 0 : true; }();
  return on;
}

// layout of a plan's device slab (offsets in bytes); also what zquatev_b200_workspace_query reports
struct Slab {
  size_t pan, x, p, cnt, pd, pt, dW, dV, np, gp, d, e, tau, al, G, L, R, P, T, T12, S12, ST, Y, TY, YP, s, bis, info, eig, scl, dcv;
  size_t bytes, yp_elems;
};
static Slab slab_layout(int n, int nb) {
  Slab o{};
  const size_t N = (size_t)n;
  size_t bytes = 0;
  auto take = [&](size_t b) { size_t at = bytes; bytes += (b + 255) & ~(size_t)255; return at; };
  o.pan = take(4 * (size_t)nb * N * sizeof(cplx));
  o.x = take((N + 3) * sizeof(quat)); o.p = take(N * sizeof(quat)); o.cnt = take(256);
  o.pd = take((size_t)cdiv(n, MV_TC) * N * sizeof(quat));
  o.pt = take((size_t)cdiv(n, MV_TR) * N * sizeof(quat));
  const size_t nch = DOT_MAX_CHUNKS + 1;
  o.dW = take(nch * nb * sizeof(quat)); o.dV = take(nch * nb * sizeof(quat));
  const size_t nparts = (size_t)cdiv(n, PANEL_ROWS) + 1;
  o.np = take(nparts * 8); o.gp = take(nparts * 8);
  o.d = take(N * 8); o.e = take(N * 8); o.tau = take(N * 8); o.al = take(N * sizeof(quat));
  o.G = take(N * nb * sizeof(quat));
  o.L = take(2 * N * 4 * nb * sizeof(cplx)); o.R = take(N * 4 * nb * sizeof(cplx));
  // P, Y, TY are sized for TWO merged panels (ZQ_BT_PAIR); T12 / S12 / ST hold the merged T factor and its cross term
  o.P = take(2 * N * 4 * nb * sizeof(cplx)); o.T = take((size_t)cdiv(n, nb) * 4 * nb * nb * sizeof(cplx));   // T of every panel
  o.T12 = take((size_t)16 * nb * nb * sizeof(cplx)); o.S12 = take((size_t)4 * nb * nb * sizeof(cplx));
  o.ST = take((size_t)4 * nb * nb * sizeof(cplx));
  o.Y = take((size_t)4 * nb * N * sizeof(cplx)); o.TY = take((size_t)4 * nb * N * sizeof(cplx));
  o.yp_elems = (size_t)YP_PARTS * 2 * nb * N;
  o.YP = take(o.yp_elems * sizeof(cplx));
  o.s = take(N * sizeof(quat)); o.bis = take((N + 8) * 8); o.info = take(256); o.eig = take(N * 8);
  o.scl = take(scale_scratch_doubles(n) * 8);
  o.dcv = take((N + 8 * PX_MAXW) * 8);
  o.bytes = bytes;
  return o;
}
static unsigned long long plan_slab_bytes(int n, int nb) { return slab_layout(n, nb).bytes; }

static int plan_create(int n, int nb, Plan** out) {
  Plan* p = new Plan();
  p->n = n;
  p->nb = nb;
  cudaError_t e0 = cudaGetDevice(&p->device);
  if (e0 != cudaSuccess) { delete p; return zq_cuda_fail(e0, __FILE__, __LINE__); }
  const Slab o = slab_layout(n, nb);
  p->yp_elems = o.yp_elems;
  cudaError_t e = cudaMalloc(&p->slab, o.bytes);
  if (e == cudaSuccess) e = small_prepare();
  if (e == cudaSuccess) e = cudaMemset(p->slab + o.cnt, 0, 256);     // arrival counter: every kernel leaves it at 0
  if (e != cudaSuccess) { delete p; return zq_cuda_fail(e, __FILE__, __LINE__); }
  char* b = p->slab;
  PanelWs& w = p->pw;
  w.n = n; w.nb = nb; w.lda = 0; w.A = nullptr; w.rank = 0; w.world = 1; w.apanel = nullptr; w.apanel_ld = 0;
  w.pan = (cplx*)(b + o.pan); w.x = (quat*)(b + o.x); w.xrec = n; w.counter = (unsigned int*)(b + o.cnt); w.p = (quat*)(b + o.p);
  w.pd = (quat*)(b + o.pd); w.pt = (quat*)(b + o.pt); w.dotW = (quat*)(b + o.dW); w.dotV = (quat*)(b + o.dV);
  w.nrm_part = (double*)(b + o.np); w.g_part = (double*)(b + o.gp);
  w.d = (double*)(b + o.d); w.e = (double*)(b + o.e); w.tau = (double*)(b + o.tau); w.alpha = (quat*)(b + o.al);
  w.G = (quat*)(b + o.G);
  p->L = (cplx*)(b + o.L); p->R = (cplx*)(b + o.R); p->P = (cplx*)(b + o.P); p->T = (cplx*)(b + o.T);
  p->T12 = (cplx*)(b + o.T12); p->S12 = (cplx*)(b + o.S12); p->ST = (cplx*)(b + o.ST);
  p->Y = (cplx*)(b + o.Y); p->TY = (cplx*)(b + o.TY); p->YP = (cplx*)(b + o.YP); p->s = (quat*)(b + o.s); p->bis = (double*)(b + o.bis);
  p->info_dev = (int*)(b + o.info); p->eig_dev = (double*)(b + o.eig); p->scl = (double*)(b + o.scl); p->dcv = (double*)(b + o.dcv);
  for (auto& ev : p->ev) cudaEventCreate(&ev);
  cudaEventCreate(&p->ev_gather);
  cudaEventCreateWithFlags(&p->done, cudaEventDisableTiming);
  for (auto& ev : p->ev_chunk) cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&p->ev_copy, cudaEventDisableTiming);
  for (auto& ev : p->ev_gath) cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
  *out = p;
  return 0;
}

static void handle_clear(Handle* h) {
  if (h->plans.empty()) return;
  int cur = -1;
  cudaGetDevice(&cur);
  if (cur != h->device) cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  for (Plan* p : h->plans) plan_free(p);
  h->plans.clear();
  h->last = nullptr;
  if (cur != h->device && cur >= 0) cudaSetDevice(cur);
}

// handle lock held.  Looks the plan up, else creates it; the least recently used plans are dropped beyond max_plans
// or when the device runs out of memory.
static int get_plan(Handle* h, int n, int nb, Plan** out) {
  for (size_t i = 0; i < h->plans.size(); ++i) {
    Plan* p = h->plans[i];
    if (p->n == n && p->nb == nb) {
      h->plans.erase(h->plans.begin() + i);
      h->plans.insert(h->plans.begin(), p);
      *out = p;
      return 0;
    }
  }
  auto drop_lru = [&]() {
    Plan* v = h->plans.back();
    h->plans.pop_back();
    if (v->done_valid) cudaEventSynchronize(v->done);
    if (h->last == v) h->last = nullptr;
    plan_free(v);
  };
  while (h->plans.size() >= h->max_plans) drop_lru();
  Plan* p = nullptr;
  int rc = plan_create(n, nb, &p);
  while (rc == -(1000 + (int)cudaErrorMemoryAllocation) && !h->plans.empty()) {
    cudaGetLastError();
    drop_lru();
    rc = plan_create(n, nb, &p);
  }
  if (rc) return rc;
  h->plans.insert(h->plans.begin(), p);
  *out = p;
  return 0;
}

static int default_handle(Handle** out) {
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return zq_cuda_fail(e, __FILE__, __LINE__);
  std::lock_guard<std::mutex> lk(g_mu);
  Handle*& h = g_default[dev & 63];
  if (!h) { h = new Handle(); h->device = dev; }
  *out = h;
  return 0;
}

// The panel (V_a,V_b,W_a,W_b: 4*nb*n complex, 67 MB at n = 16384) is re-read by col_update, the dot CTAs and
// reduce_correct of EVERY column while K1 streams gigabytes through the L2 in between.  An access-policy window
// can pin it in the 126 MB L2 (hitProp = persisting) for the duration of the reduction.  Measured on B200 at
// 2n = 32768 it is SLOWER (tridiagonalisation 6.43 -> 6.65 s: K1 loses streaming capacity), so it is off by
// default; ZQ_L2_PERSIST=1 enables it.
static void l2_window(cudaStream_t st, void* base, size_t bytes) {
  static const int on = [] { const char* e = getenv("ZQ_L2_PERSIST"); return e ? atoi(e) : 0; }();
  if (!on) return;
  static size_t max_win = 0, max_persist = 0;
  static bool init = false;
  if (!init) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceProp pr;
    if (cudaGetDeviceProperties(&pr, dev) == cudaSuccess) {
      max_win = (size_t)pr.accessPolicyMaxWindowSize;
      max_persist = (size_t)pr.persistingL2CacheMaxSize;
      if (max_persist) cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, max_persist);
    }
    cudaGetLastError();
    init = true;
  }
  if (!max_win || !max_persist) return;
  cudaStreamAttrValue v{};
  v.accessPolicyWindow.base_ptr = base;
  v.accessPolicyWindow.num_bytes = bytes < max_win ? bytes : max_win;
  v.accessPolicyWindow.hitRatio = bytes <= max_persist ? 1.0f : (float)((double)max_persist / (double)bytes);
  v.accessPolicyWindow.hitProp = bytes ? cudaAccessPropertyPersisting : cudaAccessPropertyNormal;
  v.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
  cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &v);
  if (!bytes) cudaCtxResetPersistingL2Cache();
  cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// K1-K4: reduction of (D; E) to a quaternion tridiagonal, reflectors left in the lower triangles
// ---------------------------------------------------------------------------------------------
static void tridiagonalise(Plan* p, cudaStream_t st) {
  const PanelWs& w = p->pw;
  const int n = w.n, nb = w.nb;
  cudaMemsetAsync(w.e, 0, (size_t)n * 8, st);
  cudaMemsetAsync(w.tau, 0, (size_t)n * 8, st);
  cudaMemsetAsync(w.alpha, 0, (size_t)n * sizeof(quat), st);
  if (n <= small_n_max()) {                 // K5: one CTA, one launch (small.cu)
    launch_tridiag_small(w, st);
    p->launches += 1;
    return;
  }
  if (n >= 2048) l2_window(st, w.pan, 4 * (size_t)nb * n * sizeof(cplx));
  const bool q8x = use_q8x(p);
  const bool prof = g_profile && p->timing;
  if (prof && p->k1ev.size() < 2 * (size_t)n) {
    const size_t old = p->k1ev.size();
    p->k1ev.resize(2 * (size_t)n);
    for (size_t i = old; i < p->k1ev.size(); ++i) cudaEventCreate(&p->k1ev[i]);
  }
  if (prof && p->k4ev.size() < 2 * (size_t)(n / nb + 1)) {
    const size_t old = p->k4ev.size();
    p->k4ev.resize(2 * (size_t)(n / nb + 1));
    for (size_t i = old; i < p->k4ev.size(); ++i) cudaEventCreate(&p->k4ev[i]);
  }
  for (int j0 = 0; j0 < n - 1; j0 += nb) {
    const int kb = (nb < n - 1 - j0) ? nb : n - 1 - j0;
    for (int i = 0; i < kb; ++i) {
      const int k = j0 + i;
      launch_col_update(w, k, j0, st);          // x, d_k; its last CTA forms the reflector scalars
      if (prof) cudaEventRecord(p->k1ev[2 * k], st);
      launch_matvec(w, k, j0, st);              // v = x u1^{-1} on the fly
      if (prof) cudaEventRecord(p->k1ev[2 * k + 1], st);
      launch_reduce_correct(w, k, j0, st);      // p, g; stores v
      p->launches += 3;
    }
    launch_finish_w(w, j0 + kb - 1, j0, st);
    const int r0 = j0 + kb, m = n - r0;
    if (m > 0) {
      if (use_qgemm(n)) {
        // (D + jE)[r0:, r0:] -= [V W] [W V]^H as ONE quaternion product, lower triangles
        launch_build_VW(w, r0, kb, p->L, p->R, st);
        if (prof) cudaEventRecord(p->k4ev[2 * (j0 / nb)], st);
        if (q8x)
          launch_qgemm_x(1, m, m, 2 * kb, -1.0, p->L, 2 * (size_t)m, (size_t)m, p->R, 2 * (size_t)m, (size_t)m, 1.0,
                         w.A + (size_t)r0 + (size_t)r0 * w.lda, w.lda, (size_t)n, 1, p->X8a, p->X8b, st);
        else
          launch_qgemm(0, 1, m, m, 2 * kb, -1.0, p->L, 2 * (size_t)m, (size_t)m, p->R, 2 * (size_t)m, (size_t)m, 1.0,
                       w.A + (size_t)r0 + (size_t)r0 * w.lda, w.lda, (size_t)n, 1, 1, 0, 0, 0, nullptr, st);
      } else {
        launch_build_LR(w, r0, kb, p->L, p->R, st);
        if (prof) cudaEventRecord(p->k4ev[2 * (j0 / nb)], st);
        // [D;E][r0:, r0:] -= L R^H on the lower triangles: batch 0 = D block, batch 1 = E block
        launch_zgemm(0, 1, m, m, 4 * kb, cmake(-1, 0), p->L, 2 * (size_t)m, p->R, (size_t)m, cmake(1, 0),
                     w.A + (size_t)r0 + (size_t)r0 * w.lda, w.lda, 1, 2, (size_t)m, 0, (size_t)n, st);
      }
      if (prof) cudaEventRecord(p->k4ev[2 * (j0 / nb) + 1], st);
      p->launches += 3;
    }
  }
  launch_col_update(w, n - 1, n - 1, st);   // d[n-1]
  p->launches += 1;
  if (n >= 2048) l2_window(st, w.pan, 0);
}

// Exchange-buffer layout per rank: [apanel: 2 x 64 x nmax complex][ypart: 2*PX_MAXW*nmax quats][flags: 8 + 2*rbmax*PX_MAXW u64]
static int px_rbmax(size_t nmax) { return (int)((nmax + PANEL_ROWS - 1) / PANEL_ROWS) + 1; }
static size_t px_bytes(size_t nmax) {
  return (size_t)2 * MAX_NB_PANEL * nmax * sizeof(cplx) + 2 * (size_t)PX_MAXW * nmax * sizeof(quat) + (8 + 2 * (size_t)px_rbmax(nmax) * PX_MAXW) * 8;
}

static void px_fill(PeerX& px, int g, char* base, size_t nmax) {
  px.apanel[g] = (cplx*)base;
  px.ypart[g] = (quat*)(px.apanel[g] + (size_t)2 * MAX_NB_PANEL * nmax);
  px.flags[g] = (unsigned long long*)(px.ypart[g] + 2 * (size_t)PX_MAXW * nmax);
}

static void px_teardown() {
  for (int g = 0; g < PX_MAXW; ++g)
    if (g_peer_base[g]) { cudaIpcCloseMemHandle(g_peer_base[g]); g_peer_base[g] = nullptr; }
  if (g_xb) { cudaFree(g_xb); g_xb = nullptr; }
  g_px = PeerX{};
}

// every rank allocates its buffer, the CUDA IPC handles are all-gathered over NCCL, peers are mapped
static int px_setup(int rank, int world, size_t nmax) {
  px_teardown();
  if (world > PX_MAXW) return 0;                   // fall back to NCCL collectives
  const char* off = getenv("ZQ_DIST_NCCL");
  if (off && atoi(off) != 0) return 0;
  const size_t bytes = px_bytes(nmax);
  if (cudaMalloc(&g_xb, bytes) != cudaSuccess) { cudaGetLastError(); return 0; }
  cudaMemset(g_xb, 0, bytes);
  cudaIpcMemHandle_t mine;
  if (cudaIpcGetMemHandle(&mine, g_xb) != cudaSuccess) { cudaGetLastError(); px_teardown(); return 0; }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "ipc handle size");
  char* dbuf = nullptr;
  ZQ_CUDA_CHECK(cudaMalloc(&dbuf, 64 * (size_t)world));
  ZQ_CUDA_CHECK(cudaMemcpy(dbuf + 64 * rank, &mine, 64, cudaMemcpyHostToDevice));
  ZQ_NCCL_CHECK(g_nccl.AllGather(dbuf + 64 * rank, dbuf, 64, ncclChar, g_comm, 0));
  ZQ_CUDA_CHECK(cudaStreamSynchronize(0));
  std::vector<cudaIpcMemHandle_t> all(world);
  ZQ_CUDA_CHECK(cudaMemcpy(all.data(), dbuf, 64 * (size_t)world, cudaMemcpyDeviceToHost));
  cudaFree(dbuf);
  PeerX px{};
  px.rank = rank; px.world = world; px.nmax = nmax; px.rbmax = px_rbmax(nmax);
  int ok = 1;
  for (int g = 0; g < world; ++g) {
    char* base = g_xb;
    if (g != rank) {
      void* ptr = nullptr;
      if (cudaIpcOpenMemHandle(&ptr, all[g], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; break; }
      g_peer_base[g] = ptr;
      base = (char*)ptr;
    }
    px_fill(px, g, base, nmax);
  }
  // all ranks must agree (a rank that cannot map its peers forces everyone onto the NCCL path)
  int* dflag = nullptr;
  ZQ_CUDA_CHECK(cudaMalloc(&dflag, sizeof(int)));
  ZQ_CUDA_CHECK(cudaMemcpy(dflag, &ok, sizeof(int), cudaMemcpyHostToDevice));
  ZQ_NCCL_CHECK(g_nccl.AllReduce(dflag, dflag, 1, ncclInt, ncclMin, g_comm, 0));
  ZQ_CUDA_CHECK(cudaStreamSynchronize(0));
  ZQ_CUDA_CHECK(cudaMemcpy(&ok, dflag, sizeof(int), cudaMemcpyDeviceToHost));
  cudaFree(dflag);
  if (!ok) { px_teardown(); return 0; }
  g_px = px;
  g_seq_base = 0;
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Multi-GPU reduction: D and E are distributed 1-D block-cyclic by 64-column blocks
// (every rank keeps the full array but only its own blocks are kept up to date).  Per column: the
// owner forms x and the reflector scalars and hands them to everybody, every rank multiplies its own
// column blocks (K1), the partial products are summed over the ranks, and the panel algebra is
// replicated.  The trailing update touches only the owned blocks: no exchange.
// Transport 2 (default): both exchanges are peer-memory stores issued by the panel kernels themselves
// (panel.cu), three launches per column as on one GPU.  Transport 1 (ZQ_DIST_NCCL=1 or peers that
// cannot be mapped): one ncclBroadcast of x + record and one ncclAllReduce of the partial M v per column.
// ---------------------------------------------------------------------------------------------
static int tridiagonalise_dist(Plan* p, cudaStream_t st) {
  PanelWs& w = p->pw;
  const int n = w.n, nb = w.nb, G = g_world;
  if (nb != MV_TC) return -5;                      // ownership granularity = K1 / GEMM column tile
  const bool prof = g_profile;
  if (prof && p->k1ev.size() < 2 * (size_t)n) {
    const size_t old = p->k1ev.size();
    p->k1ev.resize(2 * (size_t)n);
    for (size_t i = old; i < p->k1ev.size(); ++i) cudaEventCreate(&p->k1ev[i]);
  }
  if (prof && p->k4ev.size() < 2 * (size_t)(n / nb + 1)) {
    const size_t old = p->k4ev.size();
    p->k4ev.resize(2 * (size_t)(n / nb + 1));
    for (size_t i = old; i < p->k4ev.size(); ++i) cudaEventCreate(&p->k4ev[i]);
  }
  w.rank = g_rank;
  w.world = G;
  struct Restore {              // the cached plan must never keep the distributed geometry, whatever the exit path
    PanelWs& w;
    ~Restore() { w.rank = 0; w.world = 1; w.apanel = nullptr; w.apanel_ld = 0; }
  } restore{w};
  const bool use_px = g_px.world == G && (size_t)n <= g_px.nmax;
  if (n >= 2048) l2_window(st, w.pan, 4 * (size_t)nb * n * sizeof(cplx));
  PeerX px = g_px;
  px.info = p->info_dev;
  if (use_px) {
    w.apanel = px.apanel[g_rank];
    w.apanel_ld = px.nmax;
  } else {
    if (!p->apanel) ZQ_CUDA_CHECK(cudaMalloc(&p->apanel, (size_t)2 * MAX_NB_PANEL * n * sizeof(cplx)));
    w.apanel = p->apanel;
    w.apanel_ld = (size_t)n;
  }
  cudaMemsetAsync(w.e, 0, (size_t)n * 8, st);
  cudaMemsetAsync(w.tau, 0, (size_t)n * 8, st);
  cudaMemsetAsync(w.alpha, 0, (size_t)n * sizeof(quat), st);
  // Look-ahead of the panel exchange (peer transport; ZQ_DIST_EARLY_PUSH=0: off): the owner of the NEXT panel updates that
  // panel's 64 columns first, pushes them to every rank, and only then runs the rest of its trailing update -- the push
  // (33 MB x G over NVLink at m = 16384) leaves the critical path.  Safe without a second landing buffer: every rank has
  // read the current panel's columns before it pushed its last partial mat-vec, which the owner has already summed.
  const char* epe = getenv("ZQ_DIST_EARLY_PUSH");   // read at every solve (every rank must see the same value)
  const int early_on = epe ? atoi(epe) : 1;
  bool pre = false;                                // this panel's exchange was issued at the end of the previous panel
  for (int j0 = 0; j0 < n - 1; j0 += nb) {
    const int kb = (nb < n - 1 - j0) ? nb : n - 1 - j0;
    const int owner = (j0 / nb) % G;
    // the panel's columns (current only on their owner) go to every rank once; everything per column that depends on
    // them -- x, the reflector scalars, v -- is then formed by every rank itself
    if (use_px) {
      if (!pre) {
        const unsigned long long pseq = g_seq_base + (unsigned long long)j0 + 1ull;
        if (owner == g_rank) launch_push_panel(w, px, j0, kb, pseq, st);
        else launch_wait_panel(px, pseq, st);
        p->launches += 1;
      }
    } else {
      if (owner == g_rank) launch_pack_panel(w, p->apanel, (size_t)n, j0, kb, st);
      ZQ_NCCL_CHECK(g_nccl.Broadcast(p->apanel, p->apanel, (size_t)2 * (2 * MAX_NB_PANEL) * n, ncclDouble, owner, g_comm, st));
      p->launches += 1;
    }
    pre = false;
    for (int i = 0; i < kb; ++i) {
      const int k = j0 + i, m = n - k - 1;
      launch_col_update(w, k, j0, st);
      if (prof) cudaEventRecord(p->k1ev[2 * k], st);
      launch_matvec(w, k, j0, st);
      if (prof) cudaEventRecord(p->k1ev[2 * k + 1], st);
      if (use_px) {
        const unsigned long long seq = g_seq_base + (unsigned long long)k + 1ull;
        launch_reduce_correct_px(w, px, k, j0, seq, st);
        p->launches += 3;
      } else {
        launch_reduce_partial(w, k, st);
        ZQ_NCCL_CHECK(g_nccl.AllReduce(w.p + k + 1, w.p + k + 1, (size_t)4 * m, ncclDouble, ncclSum, g_comm, st));
        launch_correct(w, k, j0, st);
        p->launches += 5;
      }
    }
    launch_finish_w(w, j0 + kb - 1, j0, st);
    const int r0 = j0 + kb, m = n - r0;
    if (m > 0) {
      // owned 64-column blocks of the trailing matrix: global block (r0/64 + jt), jt = cb0, cb0 + G, ...
      const int b0 = r0 / MV_TC, nblk = (m + MV_TC - 1) / MV_TC;
      const int cb0 = ((g_rank - b0 % G) + G) % G;
      const int ncb = cb0 >= nblk ? 0 : (nblk - 1 - cb0) / G + 1;
      const bool q8 = use_qgemm(n);
      if (q8) launch_build_VW(w, r0, kb, p->L, p->R, st);
      else launch_build_LR(w, r0, kb, p->L, p->R, st);
      if (prof) cudaEventRecord(p->k4ev[2 * (j0 / nb)], st);
      auto k4 = [&](int c0, int nc) {              // trailing update of the owned blocks c0, c0 + G, ... (nc of them)
        if (nc <= 0) return;
        if (q8)
          launch_qgemm(0, 1, m, m, 2 * kb, -1.0, p->L, 2 * (size_t)m, (size_t)m, p->R, 2 * (size_t)m, (size_t)m, 1.0,
                       w.A + (size_t)r0 + (size_t)r0 * w.lda, w.lda, (size_t)n, 1, 1, 0, 0, 0, nullptr, st, c0, G, nc);
        else
          launch_zgemm_cb(0, 1, m, m, 4 * kb, cmake(-1, 0), p->L, 2 * (size_t)m, p->R, (size_t)m, cmake(1, 0),
                          w.A + (size_t)r0 + (size_t)r0 * w.lda, w.lda, 1, 2, (size_t)m, 0, (size_t)n, c0, G, nc, st);
      };
      const bool early = use_px && early_on && r0 < n - 1;        // a next panel exists (then kb == nb and r0 is its first column)
      if (early) {
        const int kbn = (nb < n - 1 - r0) ? nb : n - 1 - r0;
        const unsigned long long pseq = g_seq_base + (unsigned long long)r0 + 1ull;
        if (cb0 == 0) {                            // this rank owns the next panel's columns: update them, push, then the rest
          k4(0, ncb > 0 ? 1 : 0);
          launch_push_panel(w, px, r0, kbn, pseq, st);
          k4(G, ncb - 1);
        } else {
          k4(cb0, ncb);
          launch_wait_panel(px, pseq, st);
        }
        p->launches += 1;
        pre = true;
      } else {
        k4(cb0, ncb);
      }
      if (prof) cudaEventRecord(p->k4ev[2 * (j0 / nb) + 1], st);
      p->launches += 3;
    }
  }
  // d[n-1]: the last diagonal entry is current on the owner of the last column block only
  w.apanel = nullptr;
  launch_col_update(w, n - 1, n - 1, st);
  ZQ_NCCL_CHECK(g_nccl.Broadcast(w.d + n - 1, w.d + n - 1, 1, ncclDouble, ((n - 1) / nb) % G, g_comm, st));
  p->launches += 1;
  if (use_px) g_seq_base += (unsigned long long)n + 8ull;
  if (n >= 2048) l2_window(st, w.pan, 0);
  return 0;
}

// split-K geometry of Y = P^H X (M x ncols, K = m per segment, two segments): enough K-chunks for >= 8 waves
static SplitK bt_splitk(const Plan* p, int M, int ncols, int m, size_t segA, size_t segB) {
  const size_t ypart = (size_t)M * ncols;
  const int ctas1 = ((M + 63) / 64) * 2 * ((ncols + 63) / 64);
  int chunks = (8 * 444 + 2 * ctas1 - 1) / (2 * ctas1);
  const int cap_mem = (int)(p->yp_elems / ypart / 2), cap_k = m / 256;
  if (chunks > cap_mem) chunks = cap_mem;
  if (chunks > YP_MAX_CHUNKS) chunks = YP_MAX_CHUNKS;
  if (chunks > cap_k) chunks = cap_k;
  if (chunks < 1) chunks = 1;
  SplitK sk;
  sk.kc = (((m + chunks - 1) / chunks) + 7) & ~7;
  sk.chunks = (m + sk.kc - 1) / sk.kc;
  sk.segA = segA;
  sk.segB = segB;
  return sk;
}

// EXPERIMENTAL (ZQ_BT_PAIR=1, off by default; oracle: quat_kernels.backtransform_paired): panels ja < jb applied in one
// step,  H_ja H_jb = I - [Pa Pb] [[Ta, -Ta (Pa^H Pb) Tb], [0, Tb]] [Pa Pb]^H  with Pb zero-padded to the rows of Pa.
// The update GEMM then has K = 4 nb, so every C tile of X is loaded and stored half as often per flop.
static void backtransform_pair(Plan* p, cplx* X, size_t ldx, int ncols, int ja, int jb, cudaStream_t st) {
  const PanelWs& w = p->pw;
  const int n = w.n, nb = w.nb;
  const int ka = (nb < n - 1 - ja) ? nb : n - 1 - ja, kb = (nb < n - 1 - jb) ? nb : n - 1 - jb;
  const int m = n - 1 - ja, off = jb - ja;
  const int Ka = 2 * ka, Kb = 2 * kb, Kc = Ka + Kb;
  const size_t ldp = 2 * (size_t)m;
  cplx* Pa = p->P;
  cplx* Pb = p->P + (size_t)Ka * ldp;
  const cplx* Ta = p->T + (size_t)(ja / nb) * 4 * nb * nb;
  const cplx* Tb = p->T + (size_t)(jb / nb) * 4 * nb * nb;
  cplx* Xa = X + (size_t)(ja + 1);
  launch_build_phi(w, ja, ka, Pa, st);
  launch_build_phi_padded(w, jb, kb, Pb, m, off, st);
  // S = Pa^H Pb  (Ka x Kb), K = m over the a-rows and the b-rows
  {
    const SplitK sk = bt_splitk(p, Ka, Kb, m, (size_t)m, (size_t)m);
    const size_t spart = (size_t)Ka * Kb;
    launch_zgemm_splitk(1, 0, Ka, Kb, m, Pa, ldp, Pb, ldp, p->YP, (size_t)Ka, spart, 2, sk, st);
    launch_sum_parts(spart, 2 * sk.chunks, p->YP, spart, p->S12, st);
  }
  // T12 = [[Ta, -Ta S Tb], [0, Tb]]
  launch_assemble_T12(Ta, Ka, Tb, Kb, p->T12, st);
  launch_zgemm(0, 0, Ka, Kb, Kb, cmake(1, 0), p->S12, (size_t)Ka, Tb, (size_t)Kb, cmake(0, 0), p->ST, (size_t)Ka, 0, 1, 0, 0, 0, st);
  launch_zgemm(0, 0, Ka, Kb, Ka, cmake(-1, 0), Ta, (size_t)Ka, p->ST, (size_t)Ka, cmake(0, 0), p->T12 + (size_t)Ka * Kc, (size_t)Kc, 0, 1, 0, 0,
               0, st);
  // Y = [Pa Pb]^H X, TY = T12 Y, X -= [Pa Pb] TY
  {
    const SplitK sk = bt_splitk(p, Kc, ncols, m, (size_t)m, (size_t)n);
    const size_t ypart = (size_t)Kc * ncols;
    launch_zgemm_splitk(1, 0, Kc, ncols, m, p->P, ldp, Xa, ldx, p->YP, (size_t)Kc, ypart, 2, sk, st);
    launch_sum_parts(ypart, 2 * sk.chunks, p->YP, ypart, p->Y, st);
  }
  launch_zgemm(0, 0, Kc, ncols, Kc, cmake(1, 0), p->T12, (size_t)Kc, p->Y, (size_t)Kc, cmake(0, 0), p->TY, (size_t)Kc, 0, 1, 0, 0, 0, st);
  launch_zgemm(0, 0, m, ncols, Kc, cmake(-1, 0), p->P, ldp, p->TY, (size_t)Kc, cmake(1, 0), Xa, ldx, 0, 2, (size_t)m, 0, (size_t)n, st);
  p->launches += 11;
}

// split-K geometry of a quaternion product with an M x ncols result and K = m: enough pieces for >= 8 waves of 2 CTAs x 148 SMs
static SplitK q8_splitk(const Plan* p, int M, int ncols, int m) {
  const size_t ypart = (size_t)2 * M * ncols;
  const int ctas1 = ((M + 31) / 32) * ((ncols + 31) / 32);
  int chunks = (8 * 296 + ctas1 - 1) / ctas1;
  const int cap_mem = (int)(p->yp_elems / ypart), cap_k = m / 256;
  if (chunks > cap_mem) chunks = cap_mem;
  if (chunks > 2 * YP_MAX_CHUNKS) chunks = 2 * YP_MAX_CHUNKS;
  if (chunks > cap_k) chunks = cap_k;
  if (chunks < 1) chunks = 1;
  SplitK sk;
  sk.kc = (((m + chunks - 1) / chunks) + 7) & ~7;
  sk.chunks = (m + sk.kc - 1) / sk.kc;
  return sk;
}

// K6 on the quaternion GEMM with TWO panels ja < jb per step:
//   H_ja H_jb = I - [Va Vb] [[Ta, -Ta (Va^H Vb) Tb], [0, Tb]] [Va Vb]^H      (Vb zero-padded to the rows of Va),
// i.e. one merged quaternion panel of width Kq = ka + kb: Y has 128 rows and the update has K = 128, which the GEMM runs 11-14 %
// faster per flop than the single-panel shapes (profiles/r02_gemm_probe_pairs.jsonl), and X is read / written half as often.
static void backtransform_pair_q8(Plan* p, cplx* X, size_t ldx, int ncols, int ja, int jb, cudaStream_t st) {
  const PanelWs& w = p->pw;
  const int n = w.n, nb = w.nb;
  const int ka = (nb < n - 1 - ja) ? nb : n - 1 - ja, kb = (nb < n - 1 - jb) ? nb : n - 1 - jb;
  const int m = n - 1 - ja, off = jb - ja, Kq = ka + kb;
  const size_t ldp = 2 * (size_t)m;
  cplx* Vq = p->P;                                         // 2m x Kq, a-parts over b-parts
  cplx* Vqb = p->P + (size_t)ka * ldp;
  const cplx* Ta = p->T + (size_t)(ja / nb) * 4 * nb * nb;   // Phi forms; their first columns are (T_a; T_b) stacked
  const cplx* Tb = p->T + (size_t)(jb / nb) * 4 * nb * nb;
  cplx* Xa = X + (size_t)(ja + 1);
  launch_build_vq(w, ja, ka, Vq, m, 0, st);
  launch_build_vq(w, jb, kb, Vqb, m, off, st);
  // S_q = Va^H Vb (ka x kb, K = m), stacked 2ka x kb
  {
    const SplitK sk = q8_splitk(p, ka, kb, m);
    const size_t spart = (size_t)2 * ka * kb;
    launch_qgemm(1, 0, ka, kb, m, 1.0, Vq, ldp, (size_t)m, Vqb, ldp, (size_t)m, 0.0, p->YP, 2 * (size_t)ka, (size_t)ka, 0, 1, 0, 0, spart, &sk, st);
    launch_sum_parts(spart, sk.chunks, p->YP, spart, p->S12, st);
  }
  // C_q = -Ta_q (S_q Tb_q)
  launch_qgemm(0, 0, ka, kb, kb, 1.0, p->S12, 2 * (size_t)ka, (size_t)ka, Tb, 2 * (size_t)kb, (size_t)kb, 0.0, p->ST, 2 * (size_t)ka, (size_t)ka, 0, 1,
               0, 0, 0, nullptr, st);
  launch_qgemm(0, 0, ka, kb, ka, -1.0, Ta, 2 * (size_t)ka, (size_t)ka, p->ST, 2 * (size_t)ka, (size_t)ka, 0.0, p->S12, 2 * (size_t)ka, (size_t)ka, 0, 1,
               0, 0, 0, nullptr, st);
  launch_assemble_T12q(Ta, ka, Tb, kb, p->S12, p->T12, st);
  // Y = [Va Vb]^H X, TY = T12 Y, X -= [Va Vb] TY
  {
    const SplitK sk = q8_splitk(p, Kq, ncols, m);
    const size_t ypart = (size_t)2 * Kq * ncols;
    launch_qgemm(1, 0, Kq, ncols, m, 1.0, Vq, ldp, (size_t)m, Xa, ldx, (size_t)n, 0.0, p->YP, 2 * (size_t)Kq, (size_t)Kq, 0, 1, 0, 0, ypart, &sk, st);
    launch_sum_parts(ypart, sk.chunks, p->YP, ypart, p->Y, st);
  }
  launch_zgemm(0, 0, 2 * Kq, ncols, 2 * Kq, cmake(1, 0), p->T12, 2 * (size_t)Kq, p->Y, 2 * (size_t)Kq, cmake(0, 0), p->TY, 2 * (size_t)Kq, 0, 1, 0, 0, 0,
               st);
  if (use_q8x(p))
    launch_qgemm_x(0, m, ncols, Kq, -1.0, Vq, ldp, (size_t)m, p->TY, 2 * (size_t)Kq, (size_t)Kq, 1.0, Xa, ldx, (size_t)n, 0, p->X8a, p->X8b, st);
  else
    launch_qgemm(0, 0, m, ncols, Kq, -1.0, Vq, ldp, (size_t)m, p->TY, 2 * (size_t)Kq, (size_t)Kq, 1.0, Xa, ldx, (size_t)n, 0, 1, 0, 0, 0, nullptr, st);
  p->launches += 12;
}

// ---------------------------------------------------------------------------------------------
// K6: X <- H_0 ... H_{n-2} X, X = stacked (Xa; Xb), 2n x n, leading dimension ldx
// ---------------------------------------------------------------------------------------------
static void backtransform(Plan* p, cplx* X, size_t ldx, int ncols, cudaStream_t st) {
  const PanelWs& w = p->pw;
  const int n = w.n, nb = w.nb;
  if (n < 2 || ncols <= 0) return;
  const int last = ((n - 2) / nb) * nb;
  const char* pe = getenv("ZQ_BT_PAIR");                 // read at every solve (tests switch it)
  const bool q8 = use_qgemm(n);
  // two panels per step: default on the quaternion GEMM (n >= 1024), opt-in on the stacked complex path
  const bool pair = pe ? atoi(pe) != 0 : q8;
  for (int j0 = last; j0 >= 0; j0 -= nb) {
    if (pair && j0 >= nb) {                              // merge this panel with the one before it
      if (q8) backtransform_pair_q8(p, X, ldx, ncols, j0 - nb, j0, st);
      else backtransform_pair(p, X, ldx, ncols, j0 - nb, j0, st);
      j0 -= nb;
      continue;
    }
    const int kb = (nb < n - 1 - j0) ? nb : n - 1 - j0;
    const int m = n - 1 - j0;
    launch_build_phi(w, j0, kb, p->P, st);
    const size_t ldp = 2 * (size_t)m;
    cplx* Xa = X + (size_t)(j0 + 1);
    cplx* Xb = X + (size_t)(n + j0 + 1);
    if (use_qgemm(n)) {
      // quaternion form: V = (Va; Vb) is the first kb columns of P.  Y = V^H X (split-K over the m rows, parts summed in a
      // fixed order), TY = T Y (small, stacked complex), X -= V TY
      const size_t ypart = (size_t)2 * kb * ncols;
      const int ctas1 = ((kb + 31) / 32) * ((ncols + 31) / 32);
      int chunks = (8 * 296 + ctas1 - 1) / ctas1;                      // aim at >= 8 waves of 2 CTAs x 148 SMs
      const int cap_mem = (int)(p->yp_elems / ypart), cap_k = m / 256;
      if (chunks > cap_mem) chunks = cap_mem;
      if (chunks > 2 * YP_MAX_CHUNKS) chunks = 2 * YP_MAX_CHUNKS;
      if (chunks > cap_k) chunks = cap_k;
      if (chunks < 1) chunks = 1;
      SplitK sk;
      sk.kc = (((m + chunks - 1) / chunks) + 7) & ~7;
      sk.chunks = (m + sk.kc - 1) / sk.kc;
      launch_qgemm(1, 0, kb, ncols, m, 1.0, p->P, ldp, (size_t)m, Xa, ldx, (size_t)n, 0.0, p->YP, 2 * (size_t)kb, (size_t)kb, 0, 1, 0, 0,
                   ypart, &sk, st);
      launch_sum_parts(ypart, sk.chunks, p->YP, ypart, p->Y, st);
      launch_zgemm(0, 0, 2 * kb, ncols, 2 * kb, cmake(1, 0), p->T + (size_t)(j0 / nb) * 4 * nb * nb, 2 * (size_t)kb, p->Y, 2 * (size_t)kb, cmake(0, 0),
                   p->TY, 2 * (size_t)kb, 0, 1, 0, 0, 0, st);
      if (use_q8x(p))
        launch_qgemm_x(0, m, ncols, kb, -1.0, p->P, ldp, (size_t)m, p->TY, 2 * (size_t)kb, (size_t)kb, 1.0, Xa, ldx, (size_t)n, 0, p->X8a, p->X8b, st);
      else
        launch_qgemm(0, 0, m, ncols, kb, -1.0, p->P, ldp, (size_t)m, p->TY, 2 * (size_t)kb, (size_t)kb, 1.0, Xa, ldx, (size_t)n, 0, 1, 0, 0, 0,
                     nullptr, st);
      p->launches += 5;
      continue;
    }
    // Y = P^H X.  K runs over two segments (a-rows, b-rows of P and X); its 2kb x ncols output alone gives
    // 2 x ncols/32 CTAs (1024 at ncols = 16384: 2.3 waves; 128 on an 8-GPU column shard: under one wave), so the
    // segments are cut into K-chunks that run as one launch and are summed in a fixed order.
    static const int splitk = [] { const char* e = getenv("ZQ_BT_SPLITK"); return e ? atoi(e) : 1; }();
    if (!splitk) {   // development knob: the two-launch form (a-rows, then b-rows accumulated on top)
      launch_zgemm(1, 0, 2 * kb, ncols, m, cmake(1, 0), p->P, ldp, Xa, ldx, cmake(0, 0), p->Y, 2 * (size_t)kb, 0, 1, 0, 0, 0, st);
      launch_zgemm(1, 0, 2 * kb, ncols, m, cmake(1, 0), p->P + m, ldp, Xb, ldx, cmake(1, 0), p->Y, 2 * (size_t)kb, 0, 1, 0, 0, 0, st);
    } else {
      const size_t ypart = (size_t)2 * kb * ncols;
      const int ctas1 = ((2 * kb + 63) / 64) * 2 * ((ncols + 63) / 64);
      int chunks = (8 * 444 + 2 * ctas1 - 1) / (2 * ctas1);            // aim at >= 8 waves of 3 CTAs x 148 SMs
      const int cap_mem = (int)(p->yp_elems / ypart / 2), cap_k = m / 256;
      if (chunks > cap_mem) chunks = cap_mem;
      if (chunks > YP_MAX_CHUNKS) chunks = YP_MAX_CHUNKS;
      if (chunks > cap_k) chunks = cap_k;
      if (chunks < 1) chunks = 1;
      SplitK sk;
      sk.kc = (((m + chunks - 1) / chunks) + 7) & ~7;
      sk.chunks = (m + sk.kc - 1) / sk.kc;
      sk.segA = (size_t)m;                     // b-rows of P
      sk.segB = (size_t)n;                     // b-rows of X
      launch_zgemm_splitk(1, 0, 2 * kb, ncols, m, p->P, ldp, Xa, ldx, p->YP, 2 * (size_t)kb, ypart, 2, sk, st);
      launch_sum_parts(ypart, 2 * sk.chunks, p->YP, ypart, p->Y, st);
    }
    // TY = T Y
    launch_zgemm(0, 0, 2 * kb, ncols, 2 * kb, cmake(1, 0), p->T + (size_t)(j0 / nb) * 4 * nb * nb, 2 * (size_t)kb, p->Y, 2 * (size_t)kb, cmake(0, 0), p->TY,
                 2 * (size_t)kb, 0, 1, 0, 0, 0, st);
    // X -= P TY   (batch 0: a-rows, batch 1: b-rows)
    launch_zgemm(0, 0, m, ncols, 2 * kb, cmake(-1, 0), p->P, ldp, p->TY, 2 * (size_t)kb, cmake(1, 0), Xa, ldx, 0, 2,
                 (size_t)m, 0, (size_t)n, st);
    p->launches += 5;
  }
}

// full solve on device-resident operands.  Dfull: 2n x 2n complex (ld), left half = input.
// Host-pointer mode: where finished eigenvector column blocks go (downloaded on the plan's copy stream while the next
// block is back-transformed).
struct HostSink { cplx* D; size_t ld2; int host_result; };

// Eigenvector column blocks of the single-GPU host-pointer solve: a large first block, then blocks small enough that
// (a) the download of block c hides behind the back-transformation of block c+1 (PCIe moves a column ~10x faster than
// the GEMMs produce one) and (b) only the last, smaller block's download is exposed.  ZQ_E2E_CHUNKS="a,b,c" (column
// counts, the first absorbs the remainder) overrides.
static int sink_chunks(int n, int* nc) {
  int cnt = 1;
  nc[0] = n;
  if (const char* e = getenv("ZQ_E2E_CHUNKS")) {
    int v[4] = {0, 0, 0, 0}, k = 0;
    for (const char* q = e; *q && k < 4;) {
      v[k++] = atoi(q);
      while (*q && *q != ',') ++q;
      if (*q == ',') ++q;
    }
    int rest = 0;
    for (int i = 1; i < k; ++i) rest += v[i] > 0 ? v[i] : 0;
    if (k >= 1 && rest < n) {
      cnt = 0;
      nc[cnt++] = n - rest;
      for (int i = 1; i < k; ++i) if (v[i] > 0) nc[cnt++] = v[i];
    }
    return cnt;
  }
  if (n < 2048) return 1;
  // two blocks: the download of the first (7/8 of the columns) hides behind the back-transformation of the second, whose own
  // download (1/8) is what stays exposed.  (Three blocks 13/16 + 1/8 + 1/16 exposed less but cost more: the two-panel
  // back-transformation is at its best on wide blocks -- measured 2n = 32768: e2e - device 0.21 s with three blocks.)
  int last = n / 8;
  if (last < 512) last = 512;
  last = (last + 63) & ~63;
  nc[0] = n - last; nc[1] = last; cnt = 2;
  return cnt;
}

// Collective host-pointer solve (dist = 1): every rank back-transforms its shard of `per` eigenvector columns in SUB-BLOCKS; a
// finished sub-block is gathered over NVLink (stream gs) and downloaded (stream cs) while the next one is computed.  Rank 0 (or
// every rank, host_result = 0) downloads the sub-blocks of ALL `world` ranks, so a downloaded column costs it
// rho = world * t_pcie / t_backtransform of the time the GEMMs need to produce one (t_pcie ~ 20 us for the 1 MB of a column
// pair at 2n = 32768, t_backtransform ~ 190 us: rho ~ 0.1 world).  Going backwards from a small last sub-block (its download is
// what stays exposed) each earlier one may be 1/rho wider and still hide behind its successor; at most 3 sub-blocks -- every
// extra pass over the panels costs ~30 ms of per-panel fixed work at n = 16384.  The widths are the same on every rank (they
// only depend on per and world), so the NCCL calls match.  ZQ_DIST_CHUNKS="a,b,c" overrides (first absorbs the remainder);
// ZQ_DIST_PIPE=0 disables the pipeline (one block, gather, then download: the round-1 behaviour).
static int dist_sink_chunks(int per, int world, int* nc) {
  nc[0] = per; nc[1] = nc[2] = nc[3] = 0;
  if (const char* e = getenv("ZQ_DIST_PIPE")) if (atoi(e) == 0) return 1;
  if (const char* e = getenv("ZQ_DIST_CHUNKS")) {
    int v[4] = {0, 0, 0, 0}, k = 0;
    for (const char* q = e; *q && k < 4;) {
      v[k++] = atoi(q);
      while (*q && *q != ',') ++q;
      if (*q == ',') ++q;
    }
    int rest = 0, cnt = 1;
    for (int i = 1; i < k; ++i) rest += v[i] > 0 ? v[i] : 0;
    if (k >= 1 && rest < per) {
      cnt = 0;
      nc[cnt++] = per - rest;
      for (int i = 1; i < k; ++i) if (v[i] > 0) nc[cnt++] = v[i];
    }
    return cnt;
  }
  if (per < 1024) return 1;
  double rho = 0.105 * (world > 1 ? world : 1);
  if (rho > 0.9) rho = 0.9;
  int last = (per / 8 + 63) & ~63;
  if (last < 512) last = 512;
  int blk[3], cnt = 0, rem = per;
  int next = ((int)(last / rho) + 63) & ~63;
  if (rem - last > next + next / 2) {              // three sub-blocks: a middle one only when a clearly larger first one remains
    blk[cnt++] = last;
    blk[cnt++] = next;
    rem -= last + next;
  } else {                                          // two: the download of the first just hides behind the second
    last = ((int)(per * rho / (1.0 + rho)) + 63) & ~63;
    if (last < 512) last = 512;
    blk[cnt++] = last;
    rem -= last;
  }
  blk[cnt++] = rem;                                 // the first sub-block absorbs the remainder
  for (int i = 0; i < cnt; ++i) nc[i] = blk[cnt - 1 - i];
  return cnt;
}

// Back-transformation + result delivery of the collective host-pointer solve (dist = 1, host pointers, jobz = 1).
// Rank g owns eigenvector columns [g per, (g+1) per); it back-transforms them in the sub-blocks of dist_sink_chunks.  When
// sub-block c is final on the solver's stream:
//   gs : the sub-blocks c of all ranks are exchanged over NVLink -- host_result = 1: grouped ncclSend to rank 0 / ncclRecv on
//        rank 0 (only rank 0 needs them); host_result = 0: one grouped ncclBroadcast per rank -- landing at their own columns of
//        the right half of the receiver's staging array, which is free there;
//   cs : X = (U; V) of the own sub-block -> the caller's left half, Theta(X) formed in place (K10) -> the caller's right half;
//        then, once the exchange of this sub-block has arrived, the same for the other ranks' pieces (rank 0, or every rank);
// while the solver's stream back-transforms sub-block c+1.  Only the last sub-block's exchange + download stay exposed.
// The device's left half keeps the reflectors until the end, so nothing is swapped on the device (the staging array is internal).
static int deliver_dist_piped(Plan* p, cplx* Dfull, size_t ld, const double* Z, const int* perm, cudaStream_t st, const HostSink* sink) {
  const PanelWs& w = p->pw;
  const int n = w.n, G = g_world, per = (n + G - 1) / G;
  cplx* X = Dfull + (size_t)n * ld;
  int cnc[4];
  const int nchunk = dist_sink_chunks(per, G, cnc);
  const bool all = sink->host_result != 1;          // every rank wants all columns on its host
  const size_t w16 = (size_t)2 * n * sizeof(cplx);
  auto piece = [&](int r, int off, int wb, int& c0, int& nc) {   // columns of rank r's sub-block [off, off + wb) of its shard
    const int lo = r * per + off, hi = (r + 1) * per < n ? (r + 1) * per : n;
    c0 = lo < hi ? lo : hi;
    nc = (lo + wb < hi ? lo + wb : hi) - c0;
    if (nc < 0) nc = 0;
  };
  auto download_left = [&](int c0, int nc) -> int {  // on cs: X = (U; V) -> the caller's left half (reads X only)
    if (nc <= 0) return 0;
    ZQ_CUDA_CHECK(cudaMemcpy2DAsync(sink->D + (size_t)c0 * sink->ld2, sink->ld2 * sizeof(cplx), X + (size_t)c0 * ld, ld * sizeof(cplx), w16,
                                    (size_t)nc, cudaMemcpyDeviceToHost, p->cs));
    return 0;
  };
  auto download_right = [&](int c0, int nc) -> int { // on cs: X <- Theta(X) in place, -> the caller's right half
    if (nc <= 0) return 0;
    cplx* Xc = X + (size_t)c0 * ld;
    launch_theta_inplace(n, nc, Xc, ld, p->cs);
    ZQ_CUDA_CHECK(cudaMemcpy2DAsync(sink->D + (size_t)(n + c0) * sink->ld2, sink->ld2 * sizeof(cplx), Xc, ld * sizeof(cplx), w16, (size_t)nc,
                                    cudaMemcpyDeviceToHost, p->cs));
    p->launches += 1;
    return 0;
  };
  int off = 0;
  for (int c = 0; c < nchunk; ++c) {
    const int wb = cnc[c];
    int my0, mync;
    piece(g_rank, off, wb, my0, mync);
    if (mync > 0) {
      launch_scale_Z(n, mync, Z, (size_t)n, perm + my0, p->s, X + (size_t)my0 * ld, ld, st);
      backtransform(p, X + (size_t)my0 * ld, ld, mync, st);
      p->launches += 1;
    }
    if (c == nchunk - 1 && p->timing) cudaEventRecord(p->ev[4], st);
    ZQ_CUDA_CHECK(cudaEventRecord(p->ev_chunk[c], st));
    ZQ_CUDA_CHECK(cudaStreamWaitEvent(p->gs, p->ev_chunk[c], 0));
    ZQ_CUDA_CHECK(cudaStreamWaitEvent(p->cs, p->ev_chunk[c], 0));
    // exchange of the sub-blocks c (the same call sequence on every rank)
    {
      ZQ_NCCL_CHECK(g_nccl.GroupStart());
      ncclResult_t bad = ncclSuccess;               // an error inside the group must not leave it open
      for (int r = 0; r < G && bad == ncclSuccess; ++r) {
        int c0, nc;
        piece(r, off, wb, c0, nc);
        if (nc <= 0) continue;
        cplx* Xr = X + (size_t)c0 * ld;
        const size_t cnt = (size_t)2 * nc * ld;     // doubles: nc whole columns of the staging array (ld = 2n)
        if (all) {
          bad = g_nccl.Broadcast(Xr, Xr, cnt, ncclDouble, r, g_comm, p->gs);
        } else if (r != 0) {
          if (g_rank == r) bad = g_nccl.Send(Xr, cnt, ncclDouble, 0, g_comm, p->gs);
          else if (g_rank == 0) bad = g_nccl.Recv(Xr, cnt, ncclDouble, r, g_comm, p->gs);
        }
      }
      const ncclResult_t endr = g_nccl.GroupEnd();
      ZQ_NCCL_CHECK(bad);
      ZQ_NCCL_CHECK(endr);
      ZQ_CUDA_CHECK(cudaEventRecord(p->ev_gath[c], p->gs));
    }
    // downloads: X of the own piece at once; its in-place Theta only after the exchange has read it (this rank may be
    // sending it); then the other ranks' pieces
    int rc = download_left(my0, mync);
    if (rc) return rc;
    ZQ_CUDA_CHECK(cudaStreamWaitEvent(p->cs, p->ev_gath[c], 0));
    rc = download_right(my0, mync);
    if (rc) return rc;
    if (all || g_rank == 0) {
      for (int r = 0; r < G; ++r) {
        if (r == g_rank) continue;
        int c0, nc;
        piece(r, off, wb, c0, nc);
        rc = download_left(c0, nc);
        if (!rc) rc = download_right(c0, nc);
        if (rc) return rc;
      }
    }
    off += wb;
  }
  // the solver's stream continues (eigenvalues, status) after the last download and the last exchange
  ZQ_CUDA_CHECK(cudaEventRecord(p->ev_copy, p->cs));
  ZQ_CUDA_CHECK(cudaStreamWaitEvent(st, p->ev_copy, 0));
  ZQ_CUDA_CHECK(cudaStreamWaitEvent(st, p->ev_gath[nchunk - 1], 0));
  cudaError_t e2 = cudaGetLastError();
  if (e2 != cudaSuccess) return zq_cuda_fail(e2, __FILE__, __LINE__);
  return 0;
}

static int solve_device(Plan* p, cplx* Dfull, size_t ld, double* eig_dev, int jobz, int col0, int ncols, int dist, int gather, cudaStream_t st,
                        const HostSink* sink = nullptr) {
  const int n = p->n;
  PanelWs& w = p->pw;
  w.A = Dfull;
  w.lda = ld;
  w.rank = 0;
  w.world = 1;
  p->launches = 0;
  p->gather_ms = -1.0;
  zgemm_allow_3m(n >= 1024);
  cudaMemsetAsync(p->info_dev, 0, sizeof(int), st);
  if (p->timing) cudaEventRecord(p->ev[1], st);
  launch_scale_input(Dfull, ld, n, p->scl, st);     // zlascl analogue: a no-op pass unless max|a| is outside [1e-146, 1e145]
  p->launches += 3;
  {
    NvtxRange r_("zquatev_b200: tridiagonalisation (K1-K5)");
    if (dist) {
      if (!g_comm) return -6;
      const int rc = tridiagonalise_dist(p, st);
      if (rc) return rc;
    } else {
      tridiagonalise(p, st);
    }
  }
  launch_check_finite(n, w.d, w.e, p->info_dev, st);
  if (p->timing) cudaEventRecord(p->ev[2], st);
  if (!jobz) {
    if (dist && g_world > 1) {
      // K9 sharded by eigenvalue index ranges: rank g bisects indices [g per, (g+1) per), then one
      // all-gather of per doubles per rank (staged in the bisection scratch so that a ragged tail needs no padding
      // in the caller's array)
      const int per = (n + g_world - 1) / g_world;
      const int jlo = g_rank * per < n ? g_rank * per : n, jhi = (jlo + per < n) ? jlo + per : n;
      double* stage = p->dcv;                     // per * world doubles
      launch_bisect(n, w.d, w.e, stage, p->bis, st, jlo, jhi);   // writes stage[j], j in [jlo, jhi)
      ZQ_NCCL_CHECK(g_nccl.AllGather(stage + (size_t)g_rank * per, stage, (size_t)per, ncclDouble, g_comm, st));
      ZQ_CUDA_CHECK(cudaMemcpyAsync(eig_dev, stage, (size_t)n * 8, cudaMemcpyDeviceToDevice, st));
    } else {
      launch_bisect(n, w.d, w.e, eig_dev, p->bis, st);
    }
    launch_unscale_eig(n, eig_dev, p->scl, st);
    if (p->timing) cudaEventRecord(p->ev[3], st);
    if (p->timing) cudaEventRecord(p->ev[4], st);
  } else {
    if (!p->dc) {
      p->dc = dc_create(n);
      if (!p->dc) return zq_cuda_fail(cudaGetLastError(), __FILE__, __LINE__);
    }
    double* Z = nullptr;
    int* perm = nullptr;
    DcDist dd{g_rank, g_world, [](double* buf, size_t count, int rank, cudaStream_t s2) -> int {
                ncclResult_t r = g_nccl.AllGather(buf + (size_t)rank * count, buf, count, ncclDouble, g_comm, s2);
                return r == ncclSuccess ? 0 : nccl_fail(r, __LINE__);
              }};
    int rc;
    {
      NvtxRange r_("zquatev_b200: tridiagonal divide & conquer (K8)");
      rc = dc_solve(p->dc, n, w.d, w.e, eig_dev, &Z, &perm, p->info_dev, st, dist ? &dd : nullptr);
    }
    if (rc) return rc;
    launch_unscale_eig(n, eig_dev, p->scl, st);
    p->launches += dc_launches(p->dc) + 4;
    if (p->timing) cudaEventRecord(p->ev[3], st);
    cplx* X = Dfull + (size_t)n * ld;          // right half is scratch until the pairing
    NvtxRange r_bt("zquatev_b200: back-transformation + pairing (K6, K10)");
    launch_phase_chain(n, w.alpha, w.e, p->s, st);
    if (dist) {                                 // eigenvector columns split evenly over the ranks
      const int per = (n + g_world - 1) / g_world;
      col0 = g_rank * per < n ? g_rank * per : n;
      ncols = (col0 + per <= n) ? per : n - col0;
    } else if (ncols <= 0 || col0 < 0 || col0 + ncols > n) { col0 = 0; ncols = n; }
    launch_build_T_all(w, p->T, st);            // compact-WY T factor of every panel, once
    p->launches += 1;
    if (sink && dist && p->cs && p->gs) return deliver_dist_piped(p, Dfull, ld, Z, perm, st, sink);
    int cnc[4] = {ncols, 0, 0, 0};
    const int nchunk = (sink && !dist && p->cs) ? sink_chunks(ncols, cnc) : 1;
    const bool piped = sink && !dist && p->cs;
    int c0 = col0;
    for (int c = 0; c < nchunk; ++c) {
      const int nc = cnc[c];
      launch_scale_Z(n, nc, Z, (size_t)n, perm + c0, p->s, X + (size_t)c0 * ld, ld, st);
      backtransform(p, X + (size_t)c0 * ld, ld, nc, st);
      p->launches += 1;
      if (!piped) {
        launch_swap_pairing(n, nc, Dfull + (size_t)c0 * ld, ld, st);
        p->launches += 1;
      } else {
        // Host-pointer pipeline.  X = (U; V) of these columns is final and sits in the RIGHT half of the device array;
        // the left half still holds the reflectors the next blocks need, so nothing is swapped on the device.  On the
        // copy stream: X -> caller's left half, X <- Theta(X) in place (K10), X -> caller's right half -- while the
        // solver's stream back-transforms the next block.
        if (c == nchunk - 1 && p->timing) cudaEventRecord(p->ev[4], st);
        ZQ_CUDA_CHECK(cudaEventRecord(p->ev_chunk[c], st));
        ZQ_CUDA_CHECK(cudaStreamWaitEvent(p->cs, p->ev_chunk[c], 0));
        const size_t w16 = (size_t)2 * n * sizeof(cplx);
        cplx* Xc = X + (size_t)c0 * ld;
        ZQ_CUDA_CHECK(cudaMemcpy2DAsync(sink->D + (size_t)c0 * sink->ld2, sink->ld2 * sizeof(cplx), Xc, ld * sizeof(cplx), w16, (size_t)nc,
                                        cudaMemcpyDeviceToHost, p->cs));
        launch_theta_inplace(n, nc, Xc, ld, p->cs);
        ZQ_CUDA_CHECK(cudaMemcpy2DAsync(sink->D + (size_t)(n + c0) * sink->ld2, sink->ld2 * sizeof(cplx), Xc, ld * sizeof(cplx), w16, (size_t)nc,
                                        cudaMemcpyDeviceToHost, p->cs));
        p->launches += 1;
      }
      c0 += nc;
    }
    if (piped) {                                // the solver's stream continues (eigenvalues, status) after the last download
      ZQ_CUDA_CHECK(cudaEventRecord(p->ev_copy, p->cs));
      ZQ_CUDA_CHECK(cudaStreamWaitEvent(st, p->ev_copy, 0));
      cudaError_t e2 = cudaGetLastError();
      if (e2 != cudaSuccess) return zq_cuda_fail(e2, __FILE__, __LINE__);
      return 0;
    }
    if (dist && gather) {                       // every rank ends with all 2n columns
      if (p->timing) { cudaEventRecord(p->ev_gather, st); p->gather_ms = 0.0; }
      const int per = (n + g_world - 1) / g_world;
      if (n % g_world == 0) {                   // equal blocks: in-place all-gather of each half
        cplx* L = Dfull;
        cplx* R = Dfull + (size_t)n * ld;
        const size_t cnt = (size_t)2 * per * ld;
        ZQ_NCCL_CHECK(g_nccl.AllGather(L + (size_t)g_rank * per * ld, L, cnt, ncclDouble, g_comm, st));
        ZQ_NCCL_CHECK(g_nccl.AllGather(R + (size_t)g_rank * per * ld, R, cnt, ncclDouble, g_comm, st));
      } else {
        ZQ_NCCL_CHECK(g_nccl.GroupStart());
        ncclResult_t bad = ncclSuccess;         // an error inside the group must not leave it open
        for (int r = 0; r < g_world && bad == ncclSuccess; ++r) {
          const int c0 = r * per < n ? r * per : n, nc = (c0 + per <= n) ? per : n - c0;
          if (nc <= 0) continue;
          cplx* L = Dfull + (size_t)c0 * ld;
          cplx* R = Dfull + (size_t)(n + c0) * ld;
          bad = g_nccl.Broadcast(L, L, (size_t)2 * nc * ld, ncclDouble, r, g_comm, st);
          if (bad == ncclSuccess) bad = g_nccl.Broadcast(R, R, (size_t)2 * nc * ld, ncclDouble, r, g_comm, st);
        }
        const ncclResult_t endr = g_nccl.GroupEnd();
        ZQ_NCCL_CHECK(bad);
        ZQ_NCCL_CHECK(endr);
      }
    }
    if (p->timing) cudaEventRecord(p->ev[4], st);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return zq_cuda_fail(e, __FILE__, __LINE__);
  return 0;
}

static void collect_phases(Plan* p, bool host_mode) {
  float t;
  double* ms = p->phase_ms;
  for (int i = 0; i < 8; ++i) ms[i] = 0;
  if (host_mode && cudaEventElapsedTime(&t, p->ev[0], p->ev[1]) == cudaSuccess) ms[0] = t;
  if (cudaEventElapsedTime(&t, p->ev[1], p->ev[2]) == cudaSuccess) ms[1] = t;
  if (cudaEventElapsedTime(&t, p->ev[2], p->ev[3]) == cudaSuccess) ms[2] = t;
  if (cudaEventElapsedTime(&t, p->ev[3], p->ev[4]) == cudaSuccess) ms[3] = t;
  if (host_mode && cudaEventElapsedTime(&t, p->ev[4], p->ev[5]) == cudaSuccess) ms[4] = t;
  if (cudaEventElapsedTime(&t, p->ev[1], p->ev[4]) == cudaSuccess) ms[5] = t;
  if (g_profile && p->k1ev.size() >= 2 * (size_t)p->n) {
    double tot = 0;
    for (int k = 0; k + 1 < p->n; ++k)
      if (cudaEventElapsedTime(&t, p->k1ev[2 * k], p->k1ev[2 * k + 1]) == cudaSuccess) tot += t;
    ms[6] = tot;
    // trailing-update GEMMs (single-GPU path): panels [0, n-1) step nb, the last one has no trailing matrix
    double t4 = 0;
    const int nb = p->nb, n = p->n;
    for (int j0 = 0; j0 < n - 1; j0 += nb) {
      const int kb = (nb < n - 1 - j0) ? nb : n - 1 - j0;
      if (n - (j0 + kb) <= 0 || 2 * (size_t)(j0 / nb) + 1 >= p->k4ev.size()) continue;
      if (cudaEventElapsedTime(&t, p->k4ev[2 * (j0 / nb)], p->k4ev[2 * (j0 / nb) + 1]) == cudaSuccess) t4 += t;
    }
    cudaGetLastError();
    p->k4_ms = t4;
  }
  if (p->gather_ms >= 0.0) {                  // the last solve gathered: time from the end of the local pairing to ev[4]
    p->gather_ms = (cudaEventElapsedTime(&t, p->ev_gather, p->ev[4]) == cudaSuccess) ? t : 0.0;
    cudaGetLastError();
  } else {
    p->gather_ms = 0.0;
  }
  ms[7] = (double)p->launches;
}

static int check_args(int n2, void* D, int ld2, double* eig) {
  if (n2 < 0 || (n2 & 1)) return -1;
  if (!D) return -2;
  if (ld2 < n2) return -3;
  if (!eig) return -4;
  return 0;
}

// H2D of what the solver reads: only the LOWER triangles of A (Hermitian) and B (antisymmetric), column block by
// column block -- half the bytes of the left half (the reference needs both triangles; for a valid
// quaternion-Hermitian input they carry the same information).
constexpr int UP_BLK = 256;
// columns [ca, cb) (ca a multiple of UP_BLK)
static int upload_lower_range(cplx* Dfull, size_t ld, const cplx* D, size_t ld2, int n, int ca, int cb, cudaStream_t st) {
  for (int c0 = ca; c0 < cb; c0 += UP_BLK) {
    const int nc = (UP_BLK < cb - c0) ? UP_BLK : cb - c0;
    const size_t rows = (size_t)(n - c0);
    ZQ_CUDA_CHECK(cudaMemcpy2DAsync(Dfull + c0 + (size_t)c0 * ld, ld * sizeof(cplx), D + c0 + (size_t)c0 * ld2, ld2 * sizeof(cplx),
                                    rows * sizeof(cplx), (size_t)nc, cudaMemcpyHostToDevice, st));
    ZQ_CUDA_CHECK(cudaMemcpy2DAsync(Dfull + n + c0 + (size_t)c0 * ld, ld * sizeof(cplx), D + n + c0 + (size_t)c0 * ld2, ld2 * sizeof(cplx),
                                    rows * sizeof(cplx), (size_t)nc, cudaMemcpyHostToDevice, st));
  }
  return 0;
}
static int upload_lower(cplx* Dfull, size_t ld, const cplx* D, size_t ld2, int n, cudaStream_t st) {
  if (n < 1024) {               // small: one strided copy of the left half
    ZQ_CUDA_CHECK(cudaMemcpy2DAsync(Dfull, ld * sizeof(cplx), D, ld2 * sizeof(cplx), (size_t)2 * n * sizeof(cplx), (size_t)n,
                                    cudaMemcpyHostToDevice, st));
    return 0;
  }
  return upload_lower_range(Dfull, ld, D, ld2, n, 0, n, st);
}

// Collective solve with host pointers: every rank passes the same input (its own copy), so the upload is shared -- rank g moves
// the lower triangles of the column range [b_g, b_g+1) through ITS host link, the ranges cut so that the triangle areas are
// equal (b_g = n (1 - sqrt(1 - g/G)), rounded to upload blocks), and the ranges are then exchanged over NVLink (one grouped
// ncclBroadcast per range, whole columns of the staging array).  Round 1 had rank 0 upload everything and broadcast it:
// 80 ms + 12 ms at 2n = 32768 on 8 GPUs.  ZQ_DIST_UPLOAD=0 restores that.
static void upload_bounds(int n, int G, int* b) {
  b[0] = 0;
  for (int g = 1; g < G; ++g) {
    int c = (int)((double)n * (1.0 - sqrt(1.0 - (double)g / (double)G)));
    c = ((c + UP_BLK / 2) / UP_BLK) * UP_BLK;
    if (c < b[g - 1]) c = b[g - 1];
    if (c > n) c = n;
    b[g] = c;
  }
  b[G] = n;
}
static int upload_shared(cplx* Dfull, size_t ld, const cplx* D, size_t ld2, int n, cudaStream_t st) {
  int b[PX_MAXW + 65];
  const int G = g_world;
  if (G > PX_MAXW + 64) return -5;
  upload_bounds(n, G, b);
  int rc = upload_lower_range(Dfull, ld, D, ld2, n, b[g_rank], b[g_rank + 1], st);
  if (rc) return rc;
  ZQ_NCCL_CHECK(g_nccl.GroupStart());
  ncclResult_t bad = ncclSuccess;
  for (int g = 0; g < G && bad == ncclSuccess; ++g) {
    if (b[g + 1] <= b[g]) continue;
    cplx* Cg = Dfull + (size_t)b[g] * ld;
    bad = g_nccl.Broadcast(Cg, Cg, (size_t)2 * (b[g + 1] - b[g]) * ld, ncclDouble, g, g_comm, st);
  }
  const ncclResult_t endr = g_nccl.GroupEnd();
  ZQ_NCCL_CHECK(bad);
  ZQ_NCCL_CHECK(endr);
  return 0;
}

static int plan_host_staging(Plan* p) {
  const size_t n2 = 2 * (size_t)p->n;
  if (!p->Dfull) ZQ_CUDA_CHECK(cudaMalloc(&p->Dfull, n2 * n2 * sizeof(cplx)));
  if (!p->cs) ZQ_CUDA_CHECK(cudaStreamCreateWithFlags(&p->cs, cudaStreamNonBlocking));
  if (!p->gs) ZQ_CUDA_CHECK(cudaStreamCreateWithFlags(&p->gs, cudaStreamNonBlocking));
  return 0;
}

static int solve_any(Handle* h, int n2, void* D, int ld2, double* eig, const zq_options* opt) {
  NvtxRange r_("zquatev_b200: solve");
  int rc = check_args(n2, D, ld2, eig);
  if (rc) return rc;
  if (n2 == 0) return 0;
  if (!h) {
    rc = default_handle(&h);
    if (rc) return rc;
  }
  const int n = n2 / 2;
  const int jobz = opt ? opt->jobz : 1;
  const int devp = opt ? opt->device_ptrs : 0;
  const int dist = opt ? opt->dist : 0;
  int nb = (opt && opt->nb > 0) ? opt->nb : DEFAULT_NB;
  if (nb > MAX_NB) nb = MAX_NB;
  cudaStream_t st = opt ? (cudaStream_t)opt->stream : (cudaStream_t)0;
  int dev = -1;
  ZQ_CUDA_CHECK(cudaGetDevice(&dev));
  if (dev != h->device) return -7;                 // a handle is bound to the device it was created on
  std::unique_lock<std::mutex> dlk(g_dist_mu, std::defer_lock);
  if (dist) dlk.lock();                            // one communicator per process: collective solves are serialised
  std::lock_guard<std::mutex> lk(h->mu);
  Plan* p = nullptr;
  rc = get_plan(h, n, nb, &p);
  if (rc) return rc;
  h->last = p;
  // the workspace may still be in use by an asynchronous solve enqueued on ANOTHER stream: order after it
  if (p->done_valid) ZQ_CUDA_CHECK(cudaStreamWaitEvent(st, p->done, 0));
  struct DoneMark {             // every exit path (errors included) leaves the completion event behind the enqueued work
    Plan* p; cudaStream_t st;
    ~DoneMark() { if (cudaEventRecord(p->done, st) == cudaSuccess) p->done_valid = true; else cudaGetLastError(); }
  } done_mark{p, st};
  if (devp) {
    rc = solve_device(p, (cplx*)D, (size_t)ld2, eig, jobz, opt ? opt->col0 : 0, opt ? opt->ncols : 0, dist, 1, st);
    if (rc) return rc;
    if (opt && opt->sync) {
      int info = 0;
      ZQ_CUDA_CHECK(cudaMemcpyAsync(&info, p->info_dev, sizeof(int), cudaMemcpyDeviceToHost, st));
      ZQ_CUDA_CHECK(cudaStreamSynchronize(st));
      collect_phases(p, false);
      return info;
    }
    return 0;
  }
  // host pointers: stage through a device copy of the full 2n x 2n array
  const size_t ld = (size_t)n2;
  rc = plan_host_staging(p);
  if (rc) return rc;
  cudaEventRecord(p->ev[0], st);
  const char* due = getenv("ZQ_DIST_UPLOAD");      // read at every solve (every rank must see the same value)
  const int shared_up = due ? atoi(due) : 1;
  if (dist && g_comm && g_world > 1 && n >= 1024 && shared_up) {
    rc = upload_shared(p->Dfull, ld, (const cplx*)D, (size_t)ld2, n, st);   // every rank uploads a share, NVLink carries the rest
    if (rc) return rc;
  } else {
    if (!dist || g_rank == 0 || !g_comm) {
      rc = upload_lower(p->Dfull, ld, (const cplx*)D, (size_t)ld2, n, st);
      if (rc) return rc;
    }
    if (dist && g_comm)        // rank 0 uploads once, NVLink carries the input to the others
      ZQ_NCCL_CHECK(g_nccl.Broadcast(p->Dfull, p->Dfull, (size_t)2 * n * ld, ncclDouble, 0, g_comm, st));
  }
  const HostSink sink{(cplx*)D, (size_t)ld2, (opt && dist) ? opt->host_result : 0};
  // finished eigenvector column blocks are downloaded while the next is computed: one GPU (sink_chunks) and, with
  // ZQ_DIST_PIPE != 0 (default), the collective solve (deliver_dist_piped)
  const char* dpe = getenv("ZQ_DIST_PIPE");        // read at every solve (tests switch it)
  const bool dist_piped = jobz && dist && g_comm && g_world > 1 && !(dpe && atoi(dpe) == 0);
  const bool piped = jobz && (!dist || dist_piped);
  rc = solve_device(p, p->Dfull, ld, p->eig_dev, jobz, 0, 0, dist, 1, st, piped ? &sink : nullptr);
  if (rc) return rc;
  if (jobz && !piped) {
    if (dist && opt->host_result == 1 && g_rank != 0) {
      // distributed host result: this rank downloads only its own eigenvector columns (and their Kramers partners)
      const int per = (n + g_world - 1) / g_world;
      const int c0 = g_rank * per < n ? g_rank * per : n, nc = (c0 + per <= n) ? per : n - c0;
      if (nc > 0) {
        ZQ_CUDA_CHECK(cudaMemcpy2DAsync((cplx*)D + (size_t)c0 * ld2, (size_t)ld2 * sizeof(cplx), p->Dfull + (size_t)c0 * ld, ld * sizeof(cplx),
                                        (size_t)n2 * sizeof(cplx), (size_t)nc, cudaMemcpyDeviceToHost, st));
        ZQ_CUDA_CHECK(cudaMemcpy2DAsync((cplx*)D + (size_t)(n + c0) * ld2, (size_t)ld2 * sizeof(cplx), p->Dfull + (size_t)(n + c0) * ld,
                                        ld * sizeof(cplx), (size_t)n2 * sizeof(cplx), (size_t)nc, cudaMemcpyDeviceToHost, st));
      }
    } else {
      ZQ_CUDA_CHECK(cudaMemcpy2DAsync(D, (size_t)ld2 * sizeof(cplx), p->Dfull, ld * sizeof(cplx), (size_t)n2 * sizeof(cplx),
                                      (size_t)n2, cudaMemcpyDeviceToHost, st));
    }
  }
  ZQ_CUDA_CHECK(cudaMemcpyAsync(eig, p->eig_dev, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, st));
  int info = 0;
  ZQ_CUDA_CHECK(cudaMemcpyAsync(&info, p->info_dev, sizeof(int), cudaMemcpyDeviceToHost, st));
  cudaEventRecord(p->ev[5], st);
  ZQ_CUDA_CHECK(cudaStreamSynchronize(st));
  collect_phases(p, true);
  return info;
}

}  // namespace zq

using namespace zq;

// Batched small problems (BASELINE config 5): the problems are independent, so they are pipelined
// over LANES concurrent CUDA streams, each with its own plan (workspace) and pinned staging -- the
// latency-bound kernels of different problems share the GPU.  A solve of 2n = 512 is ~1200 dependent
// launches of a few microseconds each, so the host's launch rate is the limit when it enqueues them one
// by one; every lane therefore CAPTURES its solve (H2D, all kernels, D2H) once into a CUDA graph and
// replays it for the following problems: one cudaGraphLaunch per problem.  ZQ_BATCH_GRAPH=0 keeps the
// eager path (also taken if capture or instantiation fails).  (The reference has no batched entry: its
// callers loop over zquatev().)
namespace {
struct Lane {
  Plan* p = nullptr;
  cudaStream_t st = nullptr;
  int* hinfo = nullptr;
  cplx* hbuf = nullptr;
  double* heig = nullptr;
  cudaGraphExec_t gexec = nullptr;
  int uses = 0;
  bool graph_failed = false;
};
std::vector<Lane> g_lanes;
int g_lanes_n = -1, g_lanes_dev = -1;
bool g_lanes_small = false;     // reduction baked into the lanes' graphs: K5 (true) or the K1-K4 chain
int g_batch_graph_launches = 0, g_batch_eager_solves = 0;

void lanes_free() {
  for (auto& L : g_lanes) {
    if (L.st) cudaStreamSynchronize(L.st);
    if (L.gexec) cudaGraphExecDestroy(L.gexec);
    if (L.p) plan_free(L.p);
    if (L.st) cudaStreamDestroy(L.st);
    if (L.hinfo) cudaFreeHost(L.hinfo);
    if (L.hbuf) cudaFreeHost(L.hbuf);
    if (L.heig) cudaFreeHost(L.heig);
  }
  g_lanes.clear();
  g_lanes_n = -1;
  g_lanes_dev = -1;
}

// H2D of the left half, full solve, D2H of the result / eigenvalues / status: everything a lane does for one problem
int lane_enqueue(Lane& L, int n) {
  const int n2 = 2 * n;
  ZQ_CUDA_CHECK(cudaMemcpyAsync(L.p->Dfull, L.hbuf, (size_t)n * n2 * sizeof(cplx), cudaMemcpyHostToDevice, L.st));
  const int rc = solve_device(L.p, L.p->Dfull, (size_t)n2, L.p->eig_dev, 1, 0, 0, 0, 0, L.st);
  if (rc) return rc;
  ZQ_CUDA_CHECK(cudaMemcpyAsync(L.hbuf, L.p->Dfull, (size_t)n2 * n2 * sizeof(cplx), cudaMemcpyDeviceToHost, L.st));
  ZQ_CUDA_CHECK(cudaMemcpyAsync(L.heig, L.p->eig_dev, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, L.st));
  ZQ_CUDA_CHECK(cudaMemcpyAsync(L.hinfo, L.p->info_dev, sizeof(int), cudaMemcpyDeviceToHost, L.st));
  return 0;
}

// capture one solve of this lane into an executable graph; false = keep the eager path
bool lane_capture(Lane& L, int n) {
  if (cudaStreamBeginCapture(L.st, cudaStreamCaptureModeRelaxed) != cudaSuccess) { cudaGetLastError(); return false; }
  L.p->timing = false;
  const int rc = lane_enqueue(L, n);
  L.p->timing = true;
  cudaGraph_t g = nullptr;
  const cudaError_t e = cudaStreamEndCapture(L.st, &g);
  bool ok = (rc == 0 && e == cudaSuccess && g != nullptr);
  if (ok && cudaGraphInstantiate(&L.gexec, g, 0) != cudaSuccess) { L.gexec = nullptr; ok = false; }
  if (g) cudaGraphDestroy(g);
  cudaGetLastError();
  return ok;
}
}  // namespace

extern "C" {

int zquatev_b200(int n2, void* D, int ld2, double* eig) { return solve_any(nullptr, n2, D, ld2, eig, nullptr); }

int zquatev_b200_ex(int n2, void* D, int ld2, double* eig, const zq_options* opt) {
  return solve_any(nullptr, n2, D, ld2, eig, opt);
}

// ---- the step either side of the solver in the caller's workflow: products of quaternion-structured
// matrices on the device, on the same eight-product quaternion GEMM as K4 / K6.  A structured matrix
// Phi(Q) = [[Qa, -conj Qb], [Qb, conj Qa]] is passed as its LEFT half (rows [0, r) = Qa, rows [r, 2r) = Qb), exactly the
// part ts::zquatev reads and writes (zquatev.h:40-46); the right half never has to exist (fill_pairing creates it).
int zquatev_b200_qgemm(int ta, int tb, int m, int n, int k, double alpha, const void* A, int lda, const void* B, int ldb, double beta,
                       void* C, int ldc, void* stream) {
  if (m < 0 || n < 0 || k < 0) return -3;
  if (!A || !B || !C) return -7;
  const int ra = ta ? k : m, rb = tb ? n : k;          // rows of the a-part of the stored operands
  if (lda < 2 * ra || ldb < 2 * rb || ldc < 2 * m) return -8;
  if (m == 0 || n == 0) return 0;
  launch_qgemm(ta, tb, m, n, k, alpha, (const cplx*)A, (size_t)lda, (size_t)ra, (const cplx*)B, (size_t)ldb, (size_t)rb, beta, (cplx*)C,
               (size_t)ldc, (size_t)m, 0, 1, 0, 0, 0, nullptr, (cudaStream_t)stream);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : zq_cuda_fail(e, __FILE__, __LINE__);
}

// out = X^H F X for structured F (Fock-like, 2n x 2n) and X (2n x 2m): the orthogonalisation step before the eigensolver,
// and (with zquatev_b200_qgemm(0, 0, ...)) the back-multiplication C = X C' after it.  work: 2n x m complex (ld 2n).
int zquatev_b200_congruence(int n, int m, const void* X, int ldx, const void* F, int ldf, void* out, int ldo, void* work, void* stream) {
  if (n < 0 || m < 0) return -1;
  if (!X || !F || !out || !work) return -3;
  if (ldx < 2 * n || ldf < 2 * n || ldo < 2 * m) return -4;
  if (n == 0 || m == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  // W = F X  (n x m quaternion), out = X^H W
  launch_qgemm(0, 0, n, m, n, 1.0, (const cplx*)F, (size_t)ldf, (size_t)n, (const cplx*)X, (size_t)ldx, (size_t)n, 0.0, (cplx*)work, 2 * (size_t)n,
               (size_t)n, 0, 1, 0, 0, 0, nullptr, st);
  launch_qgemm(1, 0, m, m, n, 1.0, (const cplx*)X, (size_t)ldx, (size_t)n, (const cplx*)work, 2 * (size_t)n, (size_t)n, 0.0, (cplx*)out, (size_t)ldo,
               (size_t)m, 0, 1, 0, 0, 0, nullptr, st);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : zq_cuda_fail(e, __FILE__, __LINE__);
}

// right half of a structured 2r x 2c array from its left half: columns c + j = (-conj(Qb_j); conj(Qa_j))   (zquatev.cc:93-98)
int zquatev_b200_fill_pairing(int r, int c, void* M, int ld, void* stream) {
  if (r < 0 || c < 0 || !M || ld < 2 * r) return -1;
  if (r == 0 || c == 0) return 0;
  launch_fill_pairing(r, c, (cplx*)M, (size_t)ld, (cudaStream_t)stream);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : zq_cuda_fail(e, __FILE__, __LINE__);
}

// ---- handles: explicit workspace ownership (the reference's per-call allocation is zquatev.cc:63-66) ----
int zquatev_b200_create(zq_handle_t* out) {
  if (!out) return -1;
  int dev = -1;
  ZQ_CUDA_CHECK(cudaGetDevice(&dev));
  Handle* h = new Handle();
  h->device = dev;
  *out = reinterpret_cast<zq_handle_t>(h);
  return 0;
}

int zquatev_b200_destroy(zq_handle_t hh) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h) return -1;
  {
    std::lock_guard<std::mutex> lk(h->mu);
    handle_clear(h);
  }
  delete h;
  return 0;
}

int zquatev_b200_solve(zq_handle_t hh, int n2, void* D, int ld2, double* eig, const zq_options* opt) {
  if (!hh) return -1;
  return solve_any(reinterpret_cast<Handle*>(hh), n2, D, ld2, eig, opt);
}

// Allocates (or finds) the plan of this size now, including the tridiagonal D&C workspace when eigenvectors are
// wanted and the 2n x 2n device staging + copy stream of the host-pointer mode: the first solve of that size then
// allocates nothing.
int zquatev_b200_reserve(zq_handle_t hh, int n2, const zq_options* opt) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (n2 <= 0 || (n2 & 1)) return -2;
  int rc = 0;
  if (!h) { rc = default_handle(&h); if (rc) return rc; }
  int dev = -1;
  ZQ_CUDA_CHECK(cudaGetDevice(&dev));
  if (dev != h->device) return -7;
  int nb = (opt && opt->nb > 0) ? opt->nb : DEFAULT_NB;
  if (nb > MAX_NB) nb = MAX_NB;
  std::lock_guard<std::mutex> lk(h->mu);
  Plan* p = nullptr;
  rc = get_plan(h, n2 / 2, nb, &p);
  if (rc) return rc;
  if (!opt || opt->jobz) {
    if (!p->dc) p->dc = dc_create(n2 / 2);
    if (!p->dc) return zq_cuda_fail(cudaGetLastError(), __FILE__, __LINE__);
  }
  if (!opt || !opt->device_ptrs) rc = plan_host_staging(p);
  return rc;
}

// Bytes a plan of this size holds: device workspace (panels, partial sums, GEMM operands, D&C when jobz = 1, the
// 2n x 2n staging array when the operands are host pointers).  No allocation happens here.
int zquatev_b200_workspace_query(int n2, const zq_options* opt, unsigned long long* device_bytes) {
  if (n2 < 0 || (n2 & 1) || !device_bytes) return -1;
  const int n = n2 / 2;
  int nb = (opt && opt->nb > 0) ? opt->nb : DEFAULT_NB;
  if (nb > MAX_NB) nb = MAX_NB;
  unsigned long long b = plan_slab_bytes(n, nb);
  if (!opt || opt->jobz) b += dc_bytes(n);
  if (!opt || !opt->device_ptrs) b += 16ull * (unsigned long long)n2 * (unsigned long long)n2;
  if (qgemm_x_enabled() && use_qgemm(n))                                     // operand workspaces of the ZQ_Q8X variant
    b += 2ull * qgemm_x_operand_doubles(n, 2 * MAX_NB) * sizeof(double);
  *device_bytes = b;
  return 0;
}

int zquatev_b200_handle_phases(zq_handle_t hh, double ms[8]) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h) return 0;
  std::lock_guard<std::mutex> lk(h->mu);
  if (!h->last) return 0;
  for (int i = 0; i < 8; ++i) ms[i] = h->last->phase_ms[i];
  return 1;
}

// Batched entry: see the lane machinery above.  The problems are handed out by an atomic counter to a few host
// WORKER THREADS, each driving its own subset of the lanes: staging a problem costs the host ~6 MB of memcpy
// (pageable caller memory <-> pinned lane buffers), which one thread cannot do faster than ~1.3 problems/ms --
// slower than the GPU once the small-matrix reduction is a single launch (K5).  All lanes are captured up front
// by the calling thread (after one eager problem on lane 0 has set the kernel attributes), so the workers only
// copy, replay graphs and wait on their own streams.
int zquatev_b200_batched(int batch, int n2, void* D, int ld2, long long strideD, double* eig, long long strideEig,
                         int* info) {
  if (batch < 0) return -1;
  if (batch == 0) return 0;
  int rc = check_args(n2, D, ld2, eig);
  if (rc) return rc;
  if (n2 == 0) return 0;
  const int n = n2 / 2;
  static const int use_graph = [] { const char* e = getenv("ZQ_BATCH_GRAPH"); return e ? atoi(e) : 1; }();
  static const int lanes_env = [] { const char* e = getenv("ZQ_BATCH_LANES"); return e ? atoi(e) : 0; }();
  static const int threads_env = [] { const char* e = getenv("ZQ_BATCH_THREADS"); return e ? atoi(e) : 0; }();
  // a lane of the one-CTA reduction keeps one SM busy: many lanes; the multi-kernel chain fills the GPU with fewer
  const int lanes_default = !use_graph ? 8 : (n <= small_n_max() ? 96 : 16);   // measured: profiles/r01_probe_batched.jsonl
  const int LANES = lanes_env > 0 ? (lanes_env < 128 ? lanes_env : 128) : lanes_default;
  std::lock_guard<std::mutex> lk(g_mu);
  std::vector<Lane>& lanes = g_lanes;
  const int nl = batch < LANES ? batch : LANES;
  int dev = 0;
  ZQ_CUDA_CHECK(cudaGetDevice(&dev));
  const bool small_now = n <= small_n_max();
  if (g_lanes_n != n || (int)lanes.size() < nl || g_lanes_small != small_now || g_lanes_dev != dev) {   // (graphs bake the reduction in)
    lanes_free();
    // build into a local vector and commit only when every lane is complete: a failure half-way (e.g. pinned host
    // memory exhausted) must not leave half-built lanes cached for the next call
    std::vector<Lane> fresh(nl);
    auto build = [&]() -> int {
      for (auto& L : fresh) {
        int r = plan_create(n, DEFAULT_NB, &L.p);
        if (r) return r;
        ZQ_CUDA_CHECK(cudaStreamCreateWithFlags(&L.st, cudaStreamNonBlocking));
        ZQ_CUDA_CHECK(cudaMalloc(&L.p->Dfull, (size_t)n2 * n2 * sizeof(cplx)));
        L.p->dc = dc_create(n);                  // not inside a capture later
        if (!L.p->dc) return zq_cuda_fail(cudaGetLastError(), __FILE__, __LINE__);
        ZQ_CUDA_CHECK(cudaMallocHost(&L.hinfo, sizeof(int) * 4));
        // pinned staging: a D2H copy into pageable memory would block the host and serialise the lanes
        ZQ_CUDA_CHECK(cudaMallocHost(&L.hbuf, (size_t)n2 * n2 * sizeof(cplx)));
        ZQ_CUDA_CHECK(cudaMallocHost(&L.heig, (size_t)n * sizeof(double)));
      }
      return 0;
    };
    rc = build();
    lanes.swap(fresh);                           // lanes_free() releases whatever was built
    if (rc) { lanes_free(); return rc; }
    g_lanes_n = n;
    g_lanes_small = small_now;
    g_lanes_dev = dev;
  }
  std::atomic<int> next(0), worst(0), fail(0), graph_launches(0), eager_solves(0);
  // pinned staging -> caller's arrays, after the lane's stream has drained
  auto harvest = [&](int l, int& pb) -> int {
    if (pb < 0) return 0;
    Lane& L = lanes[l];
    cudaError_t e = cudaStreamSynchronize(L.st);
    if (e != cudaSuccess) return zq_cuda_fail(e, __FILE__, __LINE__);
    const int v = L.hinfo[0];
    cplx* Dp = (cplx*)D + (size_t)pb * strideD;
    for (int c = 0; c < n2; ++c) memcpy(Dp + (size_t)c * ld2, L.hbuf + (size_t)c * n2, (size_t)n2 * sizeof(cplx));
    memcpy(eig + (size_t)pb * strideEig, L.heig, (size_t)n * sizeof(double));
    if (info) info[pb] = v;
    int zero = 0;
    if (v != 0) worst.compare_exchange_strong(zero, v);
    pb = -1;
    return 0;
  };
  auto submit = [&](int l, int b) -> int {
    Lane& L = lanes[l];
    const cplx* Db = (const cplx*)D + (size_t)b * strideD;
    for (int c = 0; c < n; ++c) memcpy(L.hbuf + (size_t)c * n2, Db + (size_t)c * ld2, (size_t)n2 * sizeof(cplx));
    if (L.gexec) {
      ZQ_CUDA_CHECK(cudaGraphLaunch(L.gexec, L.st));
      ++graph_launches;
    } else {
      const int r = lane_enqueue(L, n);
      if (r) return r;
      ++eager_solves;
    }
    ++L.uses;
    return 0;
  };
  // problem 0 runs eagerly on lane 0 (first use of every kernel: attributes are set outside any capture) ...
  int first = 0;
  if (lanes[0].uses == 0 || !use_graph) {
    int pb = 0;
    rc = submit(0, 0);
    if (rc) return rc;
    rc = harvest(0, pb);
    if (rc) return rc;
    first = 1;
  }
  // ... then every lane captures its solve once
  if (use_graph)
    for (int l = 0; l < nl; ++l)
      if (!lanes[l].gexec && !lanes[l].graph_failed) lanes[l].graph_failed = !lane_capture(lanes[l], n);
  next.store(first);
  const int hw = (int)std::thread::hardware_concurrency();
  int T = threads_env > 0 ? threads_env : 12;
  if (hw > 1 && T > hw - 1) T = hw - 1;
  if (!use_graph) T = 1;                       // eager enqueue stays on the calling thread
  if (T > nl) T = nl;
  if (T < 1) T = 1;
  auto work = [&](int t) {
    cudaSetDevice(dev);
    std::vector<int> mine, pend;
    for (int l = t; l < nl; l += T) { mine.push_back(l); pend.push_back(-1); }
    size_t cur = 0;
    int r = 0;
    while (!r && !fail.load()) {
      const int b = next.fetch_add(1);
      if (b >= batch) break;
      const size_t li = cur++ % mine.size();
      r = harvest(mine[li], pend[li]);
      if (!r) r = submit(mine[li], b);
      if (!r) pend[li] = b;
    }
    for (size_t li = 0; li < mine.size(); ++li) {
      const int r2 = harvest(mine[li], pend[li]);
      if (!r) r = r2;
    }
    int zero = 0;
    if (r) fail.compare_exchange_strong(zero, r);
  };
  if (T == 1) {
    work(0);
  } else {
    std::vector<std::thread> pool;
    for (int t = 1; t < T; ++t) pool.emplace_back(work, t);
    work(0);
    for (auto& th : pool) th.join();
  }
  g_batch_graph_launches += graph_launches.load();
  g_batch_eager_solves += eager_solves.load();
  if (fail.load()) return fail.load();
  return worst.load();
}

void zquatev_b200_batched_stats(int* graph_launches, int* eager_solves) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (graph_launches) *graph_launches = g_batch_graph_launches;
  if (eager_solves) *eager_solves = g_batch_eager_solves;
}

int zquatev_b200_dist_unique_id(void* id128) {
  int rc = nccl_load();
  if (rc) return rc;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  ncclUniqueId id;
  ZQ_NCCL_CHECK(g_nccl.GetUniqueId(&id));
  memcpy(id128, &id, 128);
  return 0;
}

int zquatev_b200_dist_init(int rank, int world, const void* id128) {
  std::lock_guard<std::mutex> lk(g_mu);
  int rc = nccl_load();
  if (rc) return rc;
  if (g_comm) { px_teardown(); g_nccl.CommDestroy(g_comm); g_comm = nullptr; }
  if (world <= 1) { g_rank = 0; g_world = 1; return 0; }
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  ZQ_NCCL_CHECK(g_nccl.CommInitRank(&g_comm, world, id, rank));
  g_rank = rank;
  g_world = world;
  const char* nm = getenv("ZQ_PX_NMAX");
  return px_setup(rank, world, nm ? (size_t)atol(nm) : (size_t)32768);
}

int zquatev_b200_dist_transport(void) { return g_comm ? (g_px.world == g_world ? 2 : 1) : 0; }

void zquatev_b200_dist_finalize(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_comm) { cudaDeviceSynchronize(); px_teardown(); g_nccl.CommDestroy(g_comm); g_comm = nullptr; }
  g_rank = 0;
  g_world = 1;
}

void zquatev_b200_release(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  for (Handle* h : g_default)
    if (h) { std::lock_guard<std::mutex> lk2(h->mu); handle_clear(h); }
  lanes_free();
}

// plan of the last solve through the default handle of the current device
static Plan* last_default_plan(std::unique_lock<std::mutex>& hold) {
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  Handle* h = nullptr;
  { std::lock_guard<std::mutex> lk(g_mu); h = g_default[dev & 63]; }
  if (!h) return nullptr;
  hold = std::unique_lock<std::mutex>(h->mu);
  return h->last;
}

int zquatev_b200_last_phases(double ms[8]) {
  std::unique_lock<std::mutex> hold;
  Plan* p = last_default_plan(hold);
  if (!p) return 0;
  for (int i = 0; i < 8; ++i) ms[i] = p->phase_ms[i];
  return 1;
}

void zquatev_b200_set_profiling(int on) { g_profile = on != 0; }

double zquatev_b200_last_gather_ms(void) {
  std::unique_lock<std::mutex> hold;
  Plan* p = last_default_plan(hold);
  return p ? p->gather_ms : 0.0;
}

double zquatev_b200_last_trailing_ms(void) {
  std::unique_lock<std::mutex> hold;
  Plan* p = last_default_plan(hold);
  return p ? p->k4_ms : 0.0;
}

const char* zquatev_b200_version(void) { return "zquatev_b200 0.3 sm_100a nb=64"; }

// ---- test doors ------------------------------------------------------------------------------
int zq_test_matvec(int n, int s, const void* A, long long lda, const void* v, void* y, int reps, double* ms) {
  Handle* h = nullptr;
  int rc = default_handle(&h);
  if (rc) return rc;
  std::lock_guard<std::mutex> lk(h->mu);
  Plan* p = nullptr;
  rc = get_plan(h, n, DEFAULT_NB, &p);
  if (rc) return rc;
  PanelWs& w = p->pw;
  w.A = (cplx*)A;
  w.lda = (size_t)lda;
  cudaStream_t st = 0;
  // v itself goes into x; the record's u1^{-1} is the quaternion 1
  const double one[12] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0};
  ZQ_CUDA_CHECK(cudaMemsetAsync(w.x, 0, (size_t)n * sizeof(quat), st));
  ZQ_CUDA_CHECK(cudaMemcpyAsync(w.x + s, (const quat*)v + s, (size_t)(n - s) * sizeof(quat), cudaMemcpyDeviceToDevice, st));
  ZQ_CUDA_CHECK(cudaMemcpyAsync(w.x + w.xrec, one, sizeof(one), cudaMemcpyHostToDevice, st));
  ZQ_CUDA_CHECK(cudaStreamSynchronize(st));
  launch_matvec_only(w, s, (quat*)y, st);   // warm-up + result
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a, st);
  for (int i = 0; i < reps; ++i) launch_matvec_only(w, s, (quat*)y, st, false);   // K1 alone
  cudaEventRecord(b, st);
  ZQ_CUDA_CHECK(cudaStreamSynchronize(st));
  float t = 0;
  cudaEventElapsedTime(&t, a, b);
  if (ms) *ms = reps > 0 ? t / reps : 0.0;
  cudaEventDestroy(a); cudaEventDestroy(b);
  return 0;
}

// host-side planners of the collective host-pointer solve (pure integer logic: callable without a GPU)
int zq_test_dist_chunks(int per, int world, int out[4]) { return dist_sink_chunks(per, world, out); }
int zq_test_upload_bounds(int n, int world, int* bounds) {
  if (world < 1 || world > PX_MAXW + 64 || !bounds) return -1;
  upload_bounds(n, world, bounds);
  return 0;
}

int zq_test_zgemm(int ta, int tb, int M, int N, int K, const double* alpha, const void* A, long long lda, const void* B,
                  long long ldb, const double* beta, void* C, long long ldc, int lower, int reps, double* ms) {
  cudaStream_t st = 0;
  const cplx al = cmake(alpha[0], alpha[1]), be = cmake(beta[0], beta[1]);
  launch_zgemm(ta, tb, M, N, K, al, (const cplx*)A, (size_t)lda, (const cplx*)B, (size_t)ldb, be, (cplx*)C, (size_t)ldc,
               lower, 1, 0, 0, 0, st);
  if (reps > 0) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a, st);
    for (int i = 0; i < reps; ++i)
      launch_zgemm(ta, tb, M, N, K, al, (const cplx*)A, (size_t)lda, (const cplx*)B, (size_t)ldb, be, (cplx*)C,
                   (size_t)ldc, lower, 1, 0, 0, 0, st);
    cudaEventRecord(b, st);
    ZQ_CUDA_CHECK(cudaStreamSynchronize(st));
    float t = 0;
    cudaEventElapsedTime(&t, a, b);
    if (ms) *ms = t / reps;
    cudaEventDestroy(a); cudaEventDestroy(b);
  }
  ZQ_CUDA_CHECK(cudaStreamSynchronize(st));
  ZQ_CUDA_CHECK(cudaGetLastError());
  return 0;
}

void zq_test_set_gemm_3m(int on) { zgemm_allow_3m(on); }

int zq_test_qgemm(int ta, int tb, int M, int N, int K, double alpha, const void* A, long long lda, long long aoff, const void* B,
                  long long ldb, long long boff, double beta, void* C, long long ldc, long long coff, int lower, int reps, double* ms) {
  cudaStream_t st = 0;
  if (qgemm_x_enabled() && ta == 0 && K > 0) {       // ZQ_Q8X=1: the pre-combined-operand variant (qgemm8x.cu) behind the same door
    double *A8 = nullptr, *B8 = nullptr;
    ZQ_CUDA_CHECK(cudaMalloc(&A8, qgemm_x_operand_doubles(M, K) * sizeof(double)));
    ZQ_CUDA_CHECK(cudaMalloc(&B8, qgemm_x_operand_doubles(N, K) * sizeof(double)));
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 1 + (reps > 0 ? reps : 0); ++i) {
      if (i == 1) cudaEventRecord(a, st);
      launch_qgemm_x(tb, M, N, K, alpha, (const cplx*)A, (size_t)lda, (size_t)aoff, (const cplx*)B, (size_t)ldb, (size_t)boff, beta, (cplx*)C,
                     (size_t)ldc, (size_t)coff, lower, A8, B8, st);
    }
    cudaEventRecord(b, st);
    cudaError_t e = cudaStreamSynchronize(st);
    if (e == cudaSuccess && reps > 0 && ms) { float t = 0; cudaEventElapsedTime(&t, a, b); *ms = t / reps; }
    cudaEventDestroy(a); cudaEventDestroy(b);
    cudaFree(A8); cudaFree(B8);
    ZQ_CUDA_CHECK(e);
    ZQ_CUDA_CHECK(cudaGetLastError());
    return 0;
  }
  launch_qgemm(ta, tb, M, N, K, alpha, (const cplx*)A, (size_t)lda, (size_t)aoff, (const cplx*)B, (size_t)ldb, (size_t)boff, beta, (cplx*)C,
               (size_t)ldc, (size_t)coff, lower, 1, 0, 0, 0, nullptr, st);
  if (reps > 0) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a, st);
    for (int i = 0; i < reps; ++i)
      launch_qgemm(ta, tb, M, N, K, alpha, (const cplx*)A, (size_t)lda, (size_t)aoff, (const cplx*)B, (size_t)ldb, (size_t)boff, beta,
                   (cplx*)C, (size_t)ldc, (size_t)coff, lower, 1, 0, 0, 0, nullptr, st);
    cudaEventRecord(b, st);
    ZQ_CUDA_CHECK(cudaStreamSynchronize(st));
    float t = 0;
    cudaEventElapsedTime(&t, a, b);
    if (ms) *ms = t / reps;
    cudaEventDestroy(a); cudaEventDestroy(b);
  }
  ZQ_CUDA_CHECK(cudaStreamSynchronize(st));
  ZQ_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int zq_test_stedc(int n, const double* d, const double* e, double* w, double* Z) {
  cudaStream_t st = 0;
  DcWs* ws = dc_create(n);
  if (!ws) return zq_cuda_fail(cudaGetLastError(), __FILE__, __LINE__);
  int* info_dev = nullptr;
  cudaMalloc(&info_dev, sizeof(int));
  cudaMemsetAsync(info_dev, 0, sizeof(int), st);
  double* Zr = nullptr;
  int* perm = nullptr;
  int rc = dc_solve(ws, n, d, e, w, &Zr, &perm, info_dev, st);
  int info = 0;
  if (rc == 0) {
    // gather columns in ascending order: Z[:, j] = Zr[:, perm[j]]
    std::vector<int> hperm(n);
    cudaMemcpyAsync(hperm.data(), perm, n * sizeof(int), cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(&info, info_dev, sizeof(int), cudaMemcpyDeviceToHost, st);
    cudaStreamSynchronize(st);
    for (int j = 0; j < n; ++j)
      cudaMemcpyAsync(Z + (size_t)j * n, Zr + (size_t)hperm[j] * n, (size_t)n * 8, cudaMemcpyDeviceToDevice, st);
    cudaStreamSynchronize(st);
    cudaError_t e2 = cudaGetLastError();
    if (e2 != cudaSuccess) rc = zq_cuda_fail(e2, __FILE__, __LINE__);
  }
  cudaFree(info_dev);
  dc_destroy(ws);
  return rc ? rc : info;
}

int zq_test_bisect(int n, const double* d, const double* e, double* w) {
  cudaStream_t st = 0;
  double* scratch = nullptr;
  ZQ_CUDA_CHECK(cudaMalloc(&scratch, ((size_t)n + 8) * 8));
  launch_bisect(n, d, e, w, scratch, st);
  ZQ_CUDA_CHECK(cudaStreamSynchronize(st));
  cudaFree(scratch);
  ZQ_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int zq_test_tridiag(int n, int nb, void* A, long long lda, double* d, double* e, double* tau, double* alpha) {
  Handle* h = nullptr;
  int rc = default_handle(&h);
  if (rc) return rc;
  std::lock_guard<std::mutex> lk(h->mu);
  Plan* p = nullptr;
  if (nb <= 0) nb = DEFAULT_NB;
  if (nb > MAX_NB) nb = MAX_NB;
  rc = get_plan(h, n, nb, &p);
  if (rc) return rc;
  cudaStream_t st = 0;
  p->pw.A = (cplx*)A;
  p->pw.lda = (size_t)lda;
  tridiagonalise(p, st);
  ZQ_CUDA_CHECK(cudaMemcpyAsync(d, p->pw.d, (size_t)n * 8, cudaMemcpyDeviceToDevice, st));
  ZQ_CUDA_CHECK(cudaMemcpyAsync(e, p->pw.e, (size_t)n * 8, cudaMemcpyDeviceToDevice, st));
  ZQ_CUDA_CHECK(cudaMemcpyAsync(tau, p->pw.tau, (size_t)n * 8, cudaMemcpyDeviceToDevice, st));
  ZQ_CUDA_CHECK(cudaMemcpyAsync(alpha, p->pw.alpha, (size_t)n * sizeof(quat), cudaMemcpyDeviceToDevice, st));
  ZQ_CUDA_CHECK(cudaStreamSynchronize(st));
  ZQ_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // extern "C"

// The C++ symbol of the reference (zquatev.h:54): same mangled name, forwards to the C ABI.
namespace ts {
int zquatev(const int n2, std::complex<double>* const D, const int nld2, double* const eig) {
  return zquatev_b200(n2, static_cast<void*>(D), nld2, eig);
}
}  // namespace ts
