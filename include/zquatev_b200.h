/* zquatev_b200.h -- C ABI of the B200-native quaternionic Hermitian eigensolver.
 *
 * Drop-in boundary: the only public symbol of the reference is
 *     int ts::zquatev(const int n2, std::complex<double>* const D, const int nld2, double* const eig)
 * declared at reference zquatev.h:54 and defined at zquatev.cc:42-100.  `zquatev_b200` below has
 * exactly those arguments and semantics with plain C types; include/zquatev.h re-declares
 * ts::zquatev with the reference's signature and libzquatev_b200.so defines it by forwarding here,
 * so the reference's own test.cc links against this library unmodified.
 *
 * Everything else in this header is an EXTENSION outside the reference symbol:
 * eigenvalues-only mode, device-resident operands, batches, explicit plans, phase timings, and
 * kernel-level test doors used by tests/ (-m gpu) to compare each CUDA kernel with oracle/.
 *
 * All matrices are column-major complex<double> (interleaved re,im), as in the reference.
 * There is NO CPU fallback: every entry point needs a CUDA device (sm_100a build) and returns a
 * negative code when the CUDA runtime reports an error.
 */
#ifndef ZQUATEV_B200_H
#define ZQUATEV_B200_H

#ifdef __cplusplus
extern "C" {
#endif

/* ---- return codes (the `info` of reference zquatev.cc:84,99) -------------------------------
 *   0      success
 *   > 0    the tridiagonal eigensolver failed or the input held NaN/Inf (the reference returns
 *          zhbev's info > 0 in that case): 1 = non-finite tridiagonal,
 *          2 = secular iteration limit, 4 = leaf QL iteration limit (bits may be OR-ed)
 *   -1..-9 illegal argument number (LAPACK convention): -1 n2 odd or negative, -2 D null,
 *          -3 ld2 < n2, -4 eig null
 *   <= -1000  -(1000 + cudaError_t) CUDA runtime failure (no device, out of memory, ...)
 */

/* Replaces ts::zquatev (reference zquatev.h:54, zquatev.cc:42).
 * n2   : dimension 2n of the complex matrix (even).
 * D    : HOST array, ld2 x n2 complex<double>.  In : columns 0..n-1 hold (A; B) -- only the left
 *        half is referenced (zquatev.h:45-46) and, inside it, only the lower triangles of A
 *        (Hermitian) and B (antisymmetric).  Out: all 2n columns = ( U -V* ; V U* ).
 * ld2  : leading dimension of D (>= n2).  Unlike the reference any ld2 >= n2 works.
 * eig  : HOST array of >= n doubles; eig[0..n) ascending on return (only n values are written,
 *        as in the reference).
 */
int zquatev_b200(int n2, void* D, int ld2, double* eig);

typedef struct zq_options {
  int jobz;          /* 1: eigenvalues + eigenvectors (default), 0: eigenvalues only            */
  int device_ptrs;   /* 0: D and eig are host pointers, 1: device pointers (current device)    */
  int nb;            /* panel width of the reduction (0 = library default; reference: 20,
                        zquatev.cc:64)                                                          */
  void* stream;      /* cudaStream_t to run on (NULL = the legacy default stream)               */
  int sync;          /* device_ptrs only: 1 = wait for completion and return info (default when
                        the struct is zero-initialised is 0 = asynchronous, info not checked).  An
                        asynchronous solve keeps using the handle's plan of this size after the call
                        returns; the next solve that needs the same plan is ordered behind it on the
                        device (event wait), whatever stream it uses.                            */
  int col0, ncols;   /* eigenvector column block [col0, col0+ncols) to back-transform and return
                        (ncols = 0: all n).  Multi-GPU runs shard the back-transformation by
                        eigenvector columns: columns outside the block (and their
                        Kramers partners n+col) are left undefined.                              */
  int dist;          /* 1: collective multi-GPU solve over the communicator made by
                        zquatev_b200_dist_init: every rank calls with the SAME input (its own copy);
                        D and E are treated as distributed 1-D block-cyclic by 64-column blocks, the
                        back-transformation is split by eigenvector columns, and on return every
                        rank holds the complete result.  Needs nb = 64.                           */
  int host_result;   /* dist with host pointers: 0 = every rank downloads all 2n columns (default),
                        1 = rank 0 downloads everything, rank r > 0 only its own eigenvector columns
                        [r*ceil(n/G), ...) and their Kramers partners (the host links are shared by all
                        GPUs of a box, so G full downloads cost G times one).  Either way the ranks
                        share the upload (each moves 1/G of the lower triangles, NVLink carries the
                        rest) and finished column sub-blocks are exchanged and downloaded while the
                        next ones are back-transformed (ZQ_DIST_PIPE=0 / ZQ_DIST_UPLOAD=0: off).   */
} zq_options;

/* Same contract as zquatev_b200 plus options.  With jobz = 0 D is destroyed (holds reflectors). */
int zquatev_b200_ex(int n2, void* D, int ld2, double* eig, const zq_options* opt);

/* ---- handles -----------------------------------------------------------------
 * The reference allocates and frees its workspace inside every call (zquatev.cc:63-66, marked TODO).  A handle
 * owns a cache of plans -- the device workspace of one (n2, nb) -- on the device that was current when it was
 * created, and its own lock:
 *   - several sizes stay warm side by side (up to 4; the least recently used is dropped first, also when the
 *     device runs out of memory), so a caller alternating between sizes never re-allocates;
 *   - solves through different handles do not serialise on each other (no process-wide lock); two solves through
 *     the SAME handle are serialised on the host and, when they use different streams, ordered on the device by
 *     an event recorded at the end of every solve;
 *   - zquatev_b200 / zquatev_b200_ex / ts::zquatev use a default handle per device.
 * zquatev_b200_solve has the contract of zquatev_b200_ex.  Returns -7 when the current device is not the
 * handle's.  Collective (dist = 1) solves share the process's one communicator and are serialised.            */
typedef struct zq_handle_s* zq_handle_t;
int zquatev_b200_create(zq_handle_t* handle);
int zquatev_b200_destroy(zq_handle_t handle);
int zquatev_b200_solve(zq_handle_t handle, int n2, void* D, int ld2, double* eig, const zq_options* opt);
/* Allocates everything a solve of this size with these options needs, now (handle NULL = default handle).   */
int zquatev_b200_reserve(zq_handle_t handle, int n2, const zq_options* opt);
/* Workspace query: device bytes such a plan holds (panels, partial sums, GEMM operands; + tridiagonal D&C when
 * jobz = 1; + the 2n x 2n staging array when D is a host pointer).  Allocates nothing.                       */
int zquatev_b200_workspace_query(int n2, const zq_options* opt, unsigned long long* device_bytes);
/* phase timings of the handle's last solve (layout of zquatev_b200_last_phases)                              */
int zquatev_b200_handle_phases(zq_handle_t handle, double ms[8]);

/* ---- the steps either side of the solver in the caller ----------------------------------------
 * Products of quaternion-structured matrices Phi(Q) = ( Qa  -conj Qb ; Qb  conj Qa ) on the device, on the kernel the
 * solver's own K4 / K6 use (eight real products per quaternion product, qgemm.cu).  Every matrix is passed as its
 * LEFT half, a-part rows [0, r) stacked over b-part rows [r, 2r) -- the part ts::zquatev reads and writes
 * (reference zquatev.h:40-46); the in-tree analogue in the reference is the residual check test.cc:104-105.
 * DEVICE pointers; `stream` is a cudaStream_t (NULL = default).
 *   qgemm      : C (2m x n, ldc) = beta C + alpha op(A) op(B);  ta / tb = 1: the operand is the conjugate transpose of
 *                the stored array (stored 2k x m / 2n x k).  alpha, beta real.
 *   congruence : out (2m x m, ldo) = X^H F X  with F 2n x n (structured Fock-like matrix), X 2n x m; work >= 2n x m.
 *                The back-multiplication C = X C' after the solve is qgemm(0, 0, n, m, m, 1, X, ldx, C', ldc', 0, C, ldc).
 *   fill_pairing: columns c..2c-1 of a 2r x 2c array from columns 0..c-1 (exact: copy / conjugate / negate).   */
int zquatev_b200_qgemm(int ta, int tb, int m, int n, int k, double alpha, const void* A, int lda, const void* B, int ldb,
                       double beta, void* C, int ldc, void* stream);
int zquatev_b200_congruence(int n, int m, const void* X, int ldx, const void* F, int ldf, void* out, int ldo, void* work,
                            void* stream);
int zquatev_b200_fill_pairing(int r, int c, void* M, int ld, void* stream);

/* `batch` independent problems of the same size (BASELINE config 5): problem b uses
 * D + b*strideD (complex elements) and eig + b*strideEig; info[b] receives its return code.
 * Host pointers.  The reference has no batched entry -- its callers loop over zquatev().
 * Inside: up to 96 lanes (stream + workspace + pinned staging), each replaying a CUDA graph of its
 * whole solve; a few host worker threads stage the problems and hand them out (ZQ_BATCH_LANES,
 * ZQ_BATCH_THREADS, ZQ_BATCH_GRAPH).  For n <= 256 the reduction is ONE launch of a one-CTA kernel
 * (ZQ_SMALL_N), so the lanes run on different SMs side by side.  The call returns when every
 * problem is back in the caller's arrays.  Across GPUs the batch is sharded by the caller, one
 * contiguous shard per process (zquatev_b200/dist.py: batch_shard) -- no collective.            */
int zquatev_b200_batched(int batch, int n2, void* D, int ld2, long long strideD, double* eig,
                         long long strideEig, int* info);

/* Counters of the batched entry since the library was loaded: problems replayed from a captured CUDA
 * graph (one cudaGraphLaunch each) and problems enqueued kernel by kernel (the first use of every lane,
 * ZQ_BATCH_GRAPH=0, or a failed capture).  Either pointer may be NULL.                              */
void zquatev_b200_batched_stats(int* graph_launches, int* eager_solves);

/* Multi-GPU (one process per GPU, NCCL over NVLink).  Rank 0 creates a 128-byte NCCL
 * unique id, the caller ships it to the other ranks (e.g. torch.distributed.broadcast), then every
 * rank calls dist_init on its own CUDA device.  libnccl.so.2 is loaded lazily by these calls.   */
int zquatev_b200_dist_unique_id(void* id128);
int zquatev_b200_dist_init(int rank, int world, const void* id128);
void zquatev_b200_dist_finalize(void);
/* 0: no communicator, 1: per-column NCCL broadcast + all-reduce, 2: fused peer-memory exchange (the
 * panel kernels store into the peers' HBM over NVLink through CUDA-IPC mappings; NCCL only carries the
 * final gather).  Transport 2 is chosen when every peer can be mapped; ZQ_DIST_NCCL=1 forces 1.       */
int zquatev_b200_dist_transport(void);

/* Frees the plans of the default handles (all devices) and the lanes of the batched entry.      */
void zquatev_b200_release(void);

/* Milliseconds of the phases of the LAST call through the default handle of the current device, measured with CUDA events
 * on the solver's stream: [0] H2D, [1] tridiagonalisation, [2] tridiagonal eigensolver,
 * [3] back-transformation + pairing, [4] D2H, [5] total device time, [6] K1 mat-vec launches
 * (only when profiling is on), [7] number of kernel launches.  Returns 0 when no plan exists.  */
int zquatev_b200_last_phases(double ms[8]);
/* 1: also time every K1 launch with its own event pair (slower; for bench.py roofline).        */
void zquatev_b200_set_profiling(int on);
/* Milliseconds spent in the trailing-update GEMMs (K4) of the last PROFILED single-GPU solve (event pair
 * around each of the n/nb launches); 0 when profiling was off.                                         */
double zquatev_b200_last_trailing_ms(void);
/* Milliseconds of the NCCL gather of the eigenvector column shards at the end of the last collective
 * (dist = 1) solve -- part of phase [3]; 0 for single-GPU solves.                                       */
double zquatev_b200_last_gather_ms(void);

/* Library build info, e.g. "zquatev_b200 0.3 sm_100a nb=64".                                   */
const char* zquatev_b200_version(void);

/* ---- kernel-level test doors (device pointers; used only by tests/ and bench.py) ------------ */
/* K1: y = (D + jE) v on rows/cols [s, n) reading only the lower triangles.
 * A: 2n x n complex (lda), v,y: n quaternions (4 doubles each: a.re a.im b.re b.im), rows < s of v
 * are ignored.  Returns the average device milliseconds per launch over `reps` launches.        */
int zq_test_matvec(int n, int s, const void* A, long long lda, const void* v, void* y, int reps, double* ms);
/* K4/K6 GEMM: C = alpha op(A) op(B) + beta C (complex).                                          */
int zq_test_zgemm(int ta, int tb, int M, int N, int K, const double* alpha, const void* A, long long lda,
                  const void* B, long long ldb, const double* beta, void* C, long long ldc, int lower, int reps,
                  double* ms);
/* K4/K6 quaternion GEMM (eight real products per quaternion product): C = beta C + alpha op(A) op(B) with real alpha,
 * beta; every operand is a pair of complex arrays, the b-part `*off` elements behind the a-part; ta/tb = 1: the
 * operand is the quaternion conjugate transpose of the stored array.                                            */
int zq_test_qgemm(int ta, int tb, int M, int N, int K, double alpha, const void* A, long long lda, long long aoff,
                  const void* B, long long ldb, long long boff, double beta, void* C, long long ldc, long long coff,
                  int lower, int reps, double* ms);
/* selects the complex-product scheme of the NEXT zq_test_zgemm calls: 1 = 3M (three real products, used by the
 * solver for n >= 1024), 0 = conventional four products.  (A solve sets it again for itself.)              */
void zq_test_set_gemm_3m(int on);
/* K8: eigen-decomposition of the real symmetric tridiagonal (d, e): w[n] ascending, Z n x n (ld n). */
int zq_test_stedc(int n, const double* d, const double* e, double* w, double* Z);
/* K9: eigenvalues only by bisection.                                                             */
int zq_test_bisect(int n, const double* d, const double* e, double* w);
/* Host-side planners of the collective host-pointer solve (no device needed): widths of the sub-blocks in which every rank
 * back-transforms its `per` eigenvector columns (returns their number, <= 4; widths sum to per), and the column ranges
 * [bounds[g], bounds[g+1]) whose lower triangles rank g uploads (world + 1 entries, equal triangle areas).              */
int zq_test_dist_chunks(int per, int world, int out[4]);
int zq_test_upload_bounds(int n, int world, int* bounds);
/* K1-K4: tridiagonalise A (2n x n, lda) in place; outputs d[n], e[n], tau[n], alpha[4n].          */
int zq_test_tridiag(int n, int nb, void* A, long long lda, double* d, double* e, double* tau, double* alpha);

#ifdef __cplusplus
}
#endif
#endif
