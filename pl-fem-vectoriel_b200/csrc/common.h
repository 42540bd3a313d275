// Shared declarations of the CUDA side (context, caching device allocator, problem state).
#pragma once
#include <cuda_runtime.h>

#include <time.h>

#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/plfem.h"
#include "symbolic.h"

namespace plfem {

struct CudaError : std::runtime_error {
  using std::runtime_error::runtime_error;
};
struct StatusError : std::runtime_error {
  int status;
  StatusError(int st, const std::string& m) : std::runtime_error(m), status(st) {}
};

#define PLFEM_CUDA(call)                                                                              \
  do {                                                                                                \
    cudaError_t e_ = (call);                                                                          \
    if (e_ != cudaSuccess)                                                                            \
      throw plfem::CudaError(std::string(#call) + " failed: " + cudaGetErrorString(e_) + " at " +      \
                             __FILE__ + ":" + std::to_string(__LINE__));                              \
  } while (0)

// Wait for a stream without burning a core: spin for a few tens of microseconds (most waits in the Lanczos loop are that
// short), then poll between short naps.  A forest pool keeps several host threads per GPU waiting on their streams while
// other threads run the symbolic analysis of the next forests, and under torchrun the ranks share the node's cores:
// cudaStreamSynchronize's default busy-wait costs a core per waiting thread (measured: 6.6 ms of CPU per solve), and the
// driver's interrupt-based blocking wait (cudaDeviceScheduleBlockingSync) costs ~0.3 ms per wake-up (measured: +9 ms per
// solo solve).  PLFEM_SYNC=spin restores the plain busy-wait.
inline cudaError_t stream_wait(cudaStream_t st) {
  static const bool spin = [] { const char* e = std::getenv("PLFEM_SYNC"); return e && e[0] == 's'; }();
  if (spin) return cudaStreamSynchronize(st);
  using clk = std::chrono::steady_clock;
  const auto t0 = clk::now();
  for (;;) {
    const cudaError_t e = cudaStreamQuery(st);
    if (e != cudaErrorNotReady) return e;
    const auto waited = std::chrono::duration_cast<std::chrono::nanoseconds>(clk::now() - t0).count();
    if (waited < 40000) continue;
    // naps grow with the wait: 1/16 of it up to 0.1 ms (a solve alone: its waits last a few milliseconds); a wait that has
    // lasted 8 ms is a pool thread waiting for its forest behind the others' — 1/32 of the wait up to 0.4 ms.  A nap costs ~5 us
    // of CPU (timer, two context switches, the query): a dozen pool threads at 0.1 ms naps were half a core of a rank's four.
    const int64_t nap = waited < 8000000 ? std::min<int64_t>(100000, std::max<int64_t>(30000, waited / 16))
                                         : std::min<int64_t>(400000, waited / 32);
    timespec ts{0, (long)nap};
    nanosleep(&ts, nullptr);
  }
}

// Size-class caching allocator: cudaMalloc/cudaFree cost more than a whole assembly pass, and a
// sweep creates and drops one problem per design.
class DeviceArena {
 public:
  void* alloc(size_t bytes);
  void release(void* p);
  void destroy();
  size_t reserved_bytes() const { return reserved_; }
  size_t trim();    // cudaFree every cached (unused) block; returns the bytes given back

 private:
  bool registered_ = false;
  std::mutex mu_;   // problems of one context may be created from several host threads
  std::multimap<size_t, void*> free_;
  std::map<void*, size_t> live_;
  size_t reserved_ = 0;
};

struct SolveWork;   // device state of one solve (api.cu)

}  // namespace plfem

struct plfem_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  plfem::DeviceArena arena;
  int launches = 0;          // kernels launched since the counter was last reset
  int sweep_schedule = -1;   // PLFEM_SWEEPS_*: -1 = follow $PLFEM_SWEEP (default: dataflow), see plfem_ctx_set_sweep_schedule
  cudaEvent_t ev[8] = {};
  std::shared_ptr<plfem::SolveWork> last_work;   // what the last solve left on the device (measurement hook)
  void* pinned = nullptr;    // pinned staging buffer for small device->host reads
  size_t pinned_bytes = 0;
  void* pin(size_t bytes);
};

namespace plfem {

template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  plfem_ctx* ctx = nullptr;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { free(); }
  void free() {
    if (p && ctx) ctx->arena.release(p);
    p = nullptr; n = 0;
  }
  void alloc(plfem_ctx* c, size_t count) {
    if (p && n >= count && ctx == c) { return; }
    free();
    ctx = c; n = count;
    p = count ? static_cast<T*>(c->arena.alloc(count * sizeof(T))) : nullptr;
  }
  void upload(plfem_ctx* c, const T* h, size_t count) {
    alloc(c, count);
    if (count) PLFEM_CUDA(cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, c->stream));
  }
  void upload(plfem_ctx* c, const std::vector<T>& h) { upload(c, h.data(), h.size()); }
  void zero() {
    if (n) PLFEM_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), ctx->stream));
  }
  void download(T* h, size_t count) const {
    if (!count) return;
    PLFEM_CUDA(stream_wait(ctx->stream));     // a copy into pageable memory waits for the stream inside the driver: wait politely first
    PLFEM_CUDA(cudaMemcpyAsync(h, p, count * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
  }
};

// number of scalar value arrays produced by the two assembly modes
constexpr int NV_EXPORT = 10;  // kxx kyy kxy kyx dxx dyy dxy dyx m minv
constexpr int NV_SOLVE = 8;    // axx axy ayx ayy minv dxx dxy dyy
enum { S_AXX = 0, S_AXY, S_AYX, S_AYY, S_MINV, S_DXX, S_DXY, S_DYY };
enum { X_KXX = 0, X_KYY, X_KXY, X_KYX, X_DXX, X_DYY, X_DXY, X_DYX, X_M, X_MINV };

struct DevPattern {
  int32_t n = 0;
  int64_t nnz = 0;
  DevBuf<int32_t> rowptr, col, rowidx, old_of_new;
};

struct RefTables {  // shape functions at the quadrature points, [i*6 + q]
  double phi[36], dx[36], dy[36], w[6], qx[6], qy[6];
};
const RefTables& ref_tables();
void upload_tables();

// ---- kernels (assembly.cu) -----------------------------------------------------------------------
void launch_element_setup(plfem_ctx* ctx, const double* d_p, const int32_t* d_edofs, int64_t V, int64_t T,
                          const plfem_material& mat, const double* d_cores /* (nc,3): cx cy r */,
                          const double* d_eps_at_quad, double* d_elem);
void launch_expand_rows(plfem_ctx* ctx, const DevPattern& pat);
// device tables of one design's mesh (element DOFs, node -> elements) and its DOF count
struct PatternSource { const int32_t* n2e_ptr; const int32_t* n2e; const int32_t* edofs; int64_t N; };
// Pattern of a forest built on the device: rows node_off[b]..node_off[b+1] are design b's interior nodes in elimination
// order, old_of_new[R] the DOF id (in its own mesh) of row R; columns are global row ids.  Fills D and nnz_off[nb+1].
void build_device_pattern(plfem_ctx* ctx, int nb, const PatternSource* src, const std::vector<int32_t>& node_off,
                          const std::vector<int32_t>& old_of_new, DevPattern& D, std::vector<int64_t>& nnz_off);
void launch_assemble(plfem_ctx* ctx, const DevPattern& pat, const int32_t* d_n2e_ptr, const int32_t* d_n2e,
                     const int32_t* d_edofs, const double* d_elem, double k0sq, double alpha, int mode /* 0 solve, 1 export, 2 scalar */,
                     double* d_vals, uint32_t* d_flags);
void launch_assemble_slice(plfem_ctx* ctx, int32_t row0, int32_t nrows, const DevPattern& pat, const int32_t* d_n2e_ptr,
                           const int32_t* d_n2e, const int32_t* d_edofs, const double* d_elem, double k0sq, double alpha, int mode,
                           double* d_vals, int64_t vstride, uint32_t* d_flags);
void launch_spmv_csr(plfem_ctx* ctx, int64_t rows, const int32_t* rowptr, const int32_t* col, const double* val,
                     const double* x, double* y);

// ---- multifrontal factorisation and solves (factor.cu) ------------------------------------------------
// ---- TMA-streamed bottom subtrees of the sweeps (sweep_stream.cu) --------------------------------------------------
constexpr int ST_CHUNK_DOUBLES = 512;   // one bulk copy / ring stage: 4 KB
constexpr int ST_NLOC = 320;            // unknowns of a subtree's local vector: its pivots + the update set of its root
struct StreamSub {                      // one bottom subtree = one warp
  int64_t foff, boff;                   // chunk-aligned offsets (doubles) of its forward / backward stream
  int32_t fchunks, bchunks;
  int32_t nfronts, front0;              // its fronts (post-order) in the descriptor array
  int32_t g0, nI, next;                 // first unknown of its contiguous pivot range, pivot unknowns, unknowns of the root's update set
  int32_t uoff, soff;                   // the root's update vector in the update pool, the root's update set in strct
  int32_t pad;
};
struct StreamPackRec {                  // where one front of a subtree starts in the two streams (cursor before its items)
  int32_t f, sub, fchunk, fpos, bchunk, bpos, lm_off, pad;
};
// one task of a per-level launch (fronts above the bottom subtrees): a slab of a front's left block column (forward) or of its
// W block (backward), streamed from its own chunk-aligned piece of the level streams
struct LevelTask {
  int64_t soff, g0;                     // stream offset (doubles), first unknown of the front
  int32_t chunks, s2, u2, r0, n;        // chunks of the task's stream; pivot / update unknowns; first row (fwd) or pivot column (bwd), count
  int32_t goff;                         // forward: the front's gather tables in gsrc; backward: its update set in strct
  int32_t uoff;                         // forward: the front's update vector in the update pool
  int32_t nch;                          // forward: children of the front; backward: slabs of the front (r0 = first update unknown, n = count,
                                        // uoff = the front's partial sums, pad[0] = slab length)
  int32_t f;
  // dataflow launch (all levels in one grid): the counter this task waits on, the count it waits for, the counter it bumps
  // when its results are visible.  Forward: dep = its own front (tasks of its children above the subtrees), sig = its parent.
  // Backward: dep = its parent (all of the parent's tasks), sig = its own front.  -1 = none.
  int32_t dep, need, sig;
  int32_t pad[2];
};
struct StreamPlan {
  int n_subs = 0, n_fronts = 0;
  int64_t fwd_doubles = 0, bwd_doubles = 0;   // stream lengths including chunk padding
  DevBuf<StreamSub> subs; DevBuf<int4> fronts; DevBuf<StreamPackRec> recs; DevBuf<uint16_t> lmaps;
  DevBuf<double> sfwd, sbwd;
  // fronts above the subtrees: per-level task lists and their streams
  DevBuf<LevelTask> ftasks, btasks;
  std::vector<int32_t> fptr, bptr;            // [nlevels+1]; backward tasks are stored top level first: level l is [bptr[l+1], bptr[l])
  DevBuf<double> lfwd, lbwd;
  DevBuf<int32_t> sync;                       // counters: [0] forward ticket, [1] backward ticket, [2 + f] forward, [2 + nfronts + f] backward,
                                              // [2 + 2 nfronts + f] backward slabs of front f that have left their partial sums
  DevBuf<double> bpart;                       // partial sums of the backward slabs
  int32_t nfronts = 0;
  int64_t lfwd_doubles = 0, lbwd_doubles = 0;
};

struct DevPlan {
  int32_t n = 0, nfronts = 0, nlevels = 0;
  DevBuf<int32_t> first, s, sptr, strct, sn_of, parent, cptr, child, cmap_ptr, cmap, lfront;
  DevBuf<int64_t> foff;
  DevBuf<int32_t> uoff;              // offset of each front's update vector (in doubles)
  std::vector<int32_t> lptr;         // host copy of the level schedule
  std::vector<int32_t> lmax_m;       // largest pivot block (unknowns) per level
  std::vector<int32_t> lsplit32, lsplit64;   // per level: fronts with a pivot block <= 32 / <= 64 unknowns (listed first in lfront)
  // per-level work lists (host-built): tiles for the two GEMMs, slabs for extend-add and the sweeps
  DevBuf<int4> w_tiles, s_tiles, ea_slabs;
  DevBuf<int32_t> gsrc;
  std::vector<int32_t> w_ptr, s_ptr, ea_ptr;  // [nlevels+1] each
  StreamPlan st;                     // bottom subtrees: TMA-streamed, one warp each
  DevBuf<uint8_t> in_sub;            // front covered by a bottom subtree
  DevBuf<double> pool;               // all frontal matrices (factorisation workspace)
  DevBuf<double> upd;                // per-front update vectors of the forward sweep
  DevBuf<int32_t> status;            // [0] = 1 when a pivot block was singular
  int64_t upd_len = 0;
};
void build_dev_plan(plfem_ctx* ctx, const FrontPlan& P, DevPlan& D);
void build_stream_plan(plfem_ctx* ctx, const FrontPlan& P, const std::vector<int32_t>& uoff, std::vector<uint8_t>& in_sub, StreamPlan& S);
void build_level_plan(plfem_ctx* ctx, const FrontPlan& P, const std::vector<int32_t>& uoff, const std::vector<uint8_t>& in_sub,
                      const std::vector<int32_t>& goff, StreamPlan& S);
void launch_level_forward(plfem_ctx* ctx, const DevPlan& D, int level, const double* rhs, double* out, int nrhs, bool pdl, const uint8_t* active);
void launch_level_backward(plfem_ctx* ctx, const DevPlan& D, int level, double* x, int nrhs, bool pdl, const uint8_t* active);
// every level above the bottom subtrees in ONE launch: tasks take tickets in level order and wait on per-front counters
void reset_sweep_counters(plfem_ctx* ctx, const DevPlan& D);
void set_sweep_trace(long long* p);   // measurement only: stage clock of the dataflow forward sweep (null = off)
void launch_fused_forward(plfem_ctx* ctx, const DevPlan& D, const double* rhs, double* out, int nrhs, bool pdl, const uint8_t* active);
void launch_fused_backward(plfem_ctx* ctx, const DevPlan& D, double* x, int nrhs, bool pdl, const uint8_t* active);
void launch_stream_pack(plfem_ctx* ctx, const DevPlan& D);
void launch_stream_forward(plfem_ctx* ctx, const DevPlan& D, const double* rhs, double* out, int nrhs, bool pdl, const uint8_t* active);
void launch_stream_backward(plfem_ctx* ctx, const DevPlan& D, double* x, int nrhs, bool pdl, const uint8_t* active);
// d_sigma_node: the shift of the design each (permuted) node belongs to — a forest of designs is one problem
void launch_front_load(plfem_ctx* ctx, const DevPattern& pat, DevPlan& D, const double* d_vals, const double* d_sigma_node);
void run_factorization(plfem_ctx* ctx, const DevPlan& D);
// solves (A - sigma B) x = b in the permuted layout (Hx, Hy of a node adjacent), b and x of length 2n (must not alias);
// with nrhs = SOLVE_NRHS the right-hand sides are interleaved: entry i of right-hand side r at b[i * nrhs + r]
constexpr int SOLVE_NRHS = 4;   // block size of the multi-right-hand-side sweeps (block Lanczos)
// active_node (optional): one flag per permuted node of the forest — fronts of designs whose flag is 0 are skipped (their rows of x
// are left untouched): the refinement solve of a forest in which only some designs need it
void run_solve(plfem_ctx* ctx, const DevPlan& D, const double* b, double* x, int nrhs = 1, const uint8_t* active_node = nullptr);
void run_solve_forward(plfem_ctx* ctx, const DevPlan& D, const double* b, double* z, int nrhs = 1, const uint8_t* active_node = nullptr);
void run_solve_backward(plfem_ctx* ctx, const DevPlan& D, double* x, int nrhs = 1, bool reset_counters = true, const uint8_t* active_node = nullptr);

// ---- Lanczos + mode reductions (eigen.cu) ----------------------------------------------------------
struct EigenResult {
  std::vector<double> theta;  // Ritz values of OP, wanted ones, sorted by eigenvalue ascending
  int nconv = 0, n_op = 0, n_restart = 0, n_block_op = 0;
  int refine_steps = 0;       // refinement steps per operator application actually used
  int relaxed_from = -1;      // block step from which the relaxed operator (one refinement step fewer) was applied
  int refine_skipped = 0;     // designs of the forest that skip the refinement solve the others need
};
// A forest of independent designs laid out as one block-diagonal problem: design b owns the permuted nodes
// [noff[b], noff[b+1]) and the vector rows [moff[b], moff[b+1]) (two unknowns per node).
struct BatchDims {
  int nb = 1;
  std::vector<int32_t> noff;
  std::vector<int64_t> moff;
  DevBuf<int64_t> d_moff;
  int64_t mmax = 0;              // rows of the largest design
  void set(plfem_ctx* ctx, const std::vector<int32_t>& node_off) {
    nb = (int)node_off.size() - 1; noff = node_off; moff.resize(nb + 1); mmax = 0;
    for (int b = 0; b <= nb; ++b) moff[b] = 2 * (int64_t)noff[b];
    for (int b = 0; b < nb; ++b) mmax = std::max(mmax, moff[b + 1] - moff[b]);
    d_moff.upload(ctx, moff);
  }
};
struct DesignEig {               // one design's eigenproblem inside a forest
  int k = 0; double sigma = 0.0, tol = 1e-7;
  int status = 0; std::string err;       // PLFEM_OK or why THIS design failed (the others are unaffected)
  std::vector<double> lambda, theta;     // ascending in lambda
  int nconv = 0, n_block_op = 0, n_restart = 0;
  double solve_residual = 0.0;           // |dx| / |x| of the first refinement correction of a raw block-LDL^T solve (probe)
};
void run_eigensolver(plfem_ctx* ctx, const DevPattern& pat, DevPlan& D, const double* d_vals, const double* d_sigma_node, double sigma,
                     int k, int ncv, double tol, int maxiter, int refine_steps, const double* d_v0 /* permuted, may be null */,
                     DevBuf<double>& X /* (2n, k) eigenvectors, permuted layout */, std::vector<double>& lambda,
                     EigenResult& res);
// X is (moff[nb] x max k), leading dimension moff[nb]; per-design results and failures in `des`
void run_eigensolver_block(plfem_ctx* ctx, const DevPattern& pat, DevPlan& D, const double* d_vals, const double* d_sigma_node,
                           const BatchDims& bd, std::vector<DesignEig>& des, int ncv, int maxiter, int refine_steps,
                           const double* d_v0, DevBuf<double>& X, EigenResult& res);
void run_mode_metrics(plfem_ctx* ctx, const DevPattern& pat, const double* d_vals, const int32_t* d_perm_to_interior,
                      const uint8_t* d_in_core, const double* X, int64_t ldx /* doubles */, int32_t row0, int32_t n,
                      const std::vector<double>& lambda, int k, double* d_out_evecs /* (k, 2n) reference ordering or null */,
                      double* d_metrics /* (k,8) */, double* d_resid /* (k,2) */);

void launch_spmm_b(plfem_ctx* ctx, const DevPattern& pat, const double* d_vals, const double* x, double* y, int nrhs = 1,
                   int64_t ld = 0);
// t = b - (A - sigma B) x, x / b / t with nrhs interleaved right-hand sides
void launch_resid_k(plfem_ctx* ctx, const DevPattern& pat, const double* d_vals, const double* d_sigma_node, const double* x,
                    const double* b, double* t, int nrhs = 1, const uint8_t* active_node = nullptr);

void launch_axpy(plfem_ctx* ctx, double* x, const double* dx, int64_t m);
bool use_fused_sweeps(const plfem_ctx* ctx);   // factor.cu: dataflow launch (true) or one launch per level above the bottom subtrees
void symmetric_eigen(int n, std::vector<double>& a /* n*n col-major in, eigenvectors out */, std::vector<double>& w);
// eigenvalues + the last p rows of the eigenvector matrix only (tail[j*p + r] = Z(n-p+r, j)); a is destroyed
void symmetric_eigen_tail(int n, std::vector<double>& a, std::vector<double>& w, int p, std::vector<double>& tail);

}  // namespace plfem
