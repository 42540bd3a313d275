// C ABI entry points (include/plfem.h): context, problem, assembly/export, SpMV, modal solve.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <atomic>
#include <exception>
#include <memory>
#include <mutex>
#include <thread>

#include <malloc.h>

#include "common.h"

using namespace plfem;

// ---- host allocator --------------------------------------------------------------------------------------
// The symbolic analysis of a design builds and drops a few dozen vectors of 100 KB - 1 MB.  In a fresh C/C++ host process glibc
// serves them by mmap, then (dynamic threshold) from a heap it trims back after every analysis: ~600 page faults and 1.7 ms of
// SYSTEM time per cfg1 design, a sixth of its analysis (eight threads analysing side by side through this file's entry
// points from a C++ driver: 11.1 -> 8.9 ms of CPU per design on the build box).  With the thresholds below such blocks come
// from the heap and stay there.  A Python host has usually raised glibc's dynamic thresholds itself by the time it gets here
// (NumPy / SciPy / torch free multi-megabyte blocks while importing): there the call changes nothing measurable (0.06 ms of
// system time per design either way).  It changes the malloc parameters of the whole process; PLFEM_MALLOPT=0 leaves them alone.
static void tune_host_allocator_once() {
  static std::once_flag once;
  std::call_once(once, [] {
    const char* e = std::getenv("PLFEM_MALLOPT");
    if (e && e[0] == '0') return;
    mallopt(M_MMAP_THRESHOLD, 32 << 20);     // (the largest value glibc accepts)
    mallopt(M_TRIM_THRESHOLD, 512 << 20);
    mallopt(M_TOP_PAD, 16 << 20);
  });
}

// ---- device arena -------------------------------------------------------------------------------------
namespace plfem {

static size_t size_class(size_t bytes) {
  size_t c = 512;
  while (c < bytes) c <<= 1;
  if (c > (size_t(1) << 26)) {  // above 64 MiB: round to 32 MiB multiples instead of doubling
    const size_t g = size_t(1) << 25;
    c = (bytes + g - 1) / g * g;
  }
  return c;
}

// Every arena of the process (one per context: a forest pool keeps a dozen on one GPU).  Blocks an arena has cached but is not
// using are what fills the device when many contexts have each seen a large forest: an arena that cannot get memory first gives
// back its own cache, then asks all the others to do the same, and only then fails.
static std::mutex g_arenas_mu;
static std::vector<DeviceArena*> g_arenas;

size_t DeviceArena::trim() {
  std::lock_guard<std::mutex> g(mu_);
  size_t freed = 0;
  for (auto& kv : free_) { cudaFree(kv.second); reserved_ -= kv.first; freed += kv.first; }
  free_.clear();
  return freed;
}

void* DeviceArena::alloc(size_t bytes) {
  const size_t c = size_class(bytes);
  if (!registered_) {
    std::lock_guard<std::mutex> g(g_arenas_mu);
    if (!registered_) { g_arenas.push_back(this); registered_ = true; }
  }
  for (int attempt = 0;; ++attempt) {
    {
      std::lock_guard<std::mutex> g(mu_);
      auto it = free_.find(c);
      void* p = nullptr;
      if (it != free_.end()) {
        p = it->second;
        free_.erase(it);
        live_[p] = c;
        return p;
      }
      const cudaError_t e = cudaMalloc(&p, c);
      if (e == cudaSuccess) { reserved_ += c; live_[p] = c; return p; }
      (void)cudaGetLastError();     // the failure must not stay behind as the "last error" of a later, successful launch
      if (attempt >= 2)
        throw CudaError(std::string("cudaMalloc of ") + std::to_string(c) + " bytes failed: " + cudaGetErrorString(e));
    }
    // (no arena lock is held here: two arenas trimming each other cannot deadlock)
    if (attempt == 0) {
      trim();
    } else {
      std::lock_guard<std::mutex> g(g_arenas_mu);
      for (DeviceArena* a : g_arenas) a->trim();
    }
  }
}

void DeviceArena::release(void* p) {
  std::lock_guard<std::mutex> g(mu_);
  auto it = live_.find(p);
  if (it == live_.end()) return;
  free_.emplace(it->second, p);
  live_.erase(it);
}

void DeviceArena::destroy() {
  {
    std::lock_guard<std::mutex> g(g_arenas_mu);
    g_arenas.erase(std::remove(g_arenas.begin(), g_arenas.end(), this), g_arenas.end());
    registered_ = false;
  }
  std::lock_guard<std::mutex> g(mu_);
  for (auto& kv : free_) cudaFree(kv.second);
  for (auto& kv : live_) cudaFree(kv.first);
  free_.clear(); live_.clear(); reserved_ = 0;
}

}  // namespace plfem

void* plfem_ctx::pin(size_t bytes) {
  if (bytes > pinned_bytes) {
    if (pinned) cudaFreeHost(pinned);
    pinned = nullptr;
    size_t nb = std::max<size_t>(bytes, 1 << 20);
    PLFEM_CUDA(cudaMallocHost(&pinned, nb));
    pinned_bytes = nb;
  }
  return pinned;
}

namespace plfem {
// ---- device state of one solve: a single design, or a forest of designs as one block-diagonal problem ------
struct SolveWork {
  BatchDims bd;
  std::vector<int64_t> nnz_off;
  std::vector<int32_t> front_off;
  FrontPlan plan;                              // merged front plan
  DevPattern dpat;
  DevPlan dplan;
  DevBuf<double> d_vals;                       // NV_SOLVE x nnz
  DevBuf<double> d_sigma;                      // shift of the design each permuted node belongs to
  DevBuf<int32_t> d_perm;                      // new id -> interior index (local to its design)
  DevBuf<uint8_t> d_in_core;
  bool ready = false;
  int plan_serial = -1;                        // single design: the plan generation these copies were made from
  // what the measurement hook needs to re-run the assembly of each design (device pointers owned by the problems,
  // which must outlive the hook call)
  struct AsmArgs {
    const double* d_p; const int32_t* d_edofs; const int32_t* d_n2e_ptr; const int32_t* d_n2e; double* d_elem;
    const double* d_cores; int64_t V, T; plfem_material mat;
  };
  std::vector<AsmArgs> asm_args;
};
}  // namespace plfem

// ---- problem ------------------------------------------------------------------------------------------
struct plfem_problem {
  plfem_ctx* ctx = nullptr;
  DofTables dof;
  std::vector<double> p_host;                 // (2,V)
  DevBuf<double> d_p;                          // (2,V)
  DevBuf<int32_t> d_edofs, d_n2e_ptr, d_n2e;
  DevBuf<double> d_elem, d_cores, d_epsq;
  // full N x N pattern (export path)
  Pattern full; DevPattern dfull; bool full_ready = false;
  DevBuf<double> d_full_vals; DevBuf<uint32_t> d_full_flags;
  std::vector<double> h_full_vals; std::vector<uint32_t> h_full_flags; bool assembled = false;
  plfem_material last_mat{};
  // interior solve path (host side; the device side of a solve lives in a SolveWork)
  Pattern adj;  bool adj_ready = false;        // interior nodes, interior-index numbering (for dissection)
  FrontPlan plan; bool plan_ready = false; SymbolicOptions plan_opt;
  uint64_t mesh_print = 0;                     // fingerprint of (p, t): designs of one forest with equal meshes can share one analysis
  std::shared_ptr<SolveWork> work;             // device state of the last single-design solve (profile / debug hooks, reuse)
  int plan_serial = 0;                         // bumped whenever the front plan is rebuilt
};

namespace {

template <class F>
int guarded(plfem_ctx* ctx, F&& fn) {
  try {
    fn();
    return PLFEM_OK;
  } catch (const StatusError& e) {
    if (ctx) ctx->err = e.what();
    return e.status;
  } catch (const CudaError& e) {
    if (ctx) ctx->err = e.what();
    return PLFEM_ERR_CUDA;
  } catch (const std::bad_alloc&) {
    if (ctx) ctx->err = "host allocation failed";
    return PLFEM_ERR_INTERNAL;
  } catch (const std::exception& e) {
    if (ctx) ctx->err = e.what();
    return PLFEM_ERR_INTERNAL;
  }
}

void need(bool ok, const char* msg) {
  if (!ok) throw StatusError(PLFEM_ERR_INVALID, msg);
}

double cpu_ms() {      // CPU time of the whole process (all threads): what a forest costs the host, not how long it waits
  timespec ts;
  clock_gettime(CLOCK_PROCESS_CPUTIME_ID, &ts);
  return 1e3 * ts.tv_sec + 1e-6 * ts.tv_nsec;
}

double now_ms() {
  using namespace std::chrono;
  return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

void upload_material(plfem_problem* pb, const plfem_material* mat) {
  plfem_ctx* ctx = pb->ctx;
  need(mat != nullptr, "material is NULL");
  need(mat->n_cores >= 0, "n_cores < 0");
  need(mat->eps_at_quad || mat->n_cores == 0 || (mat->cores_xy && mat->cores_r), "core arrays are NULL");
  std::vector<double> cores(3 * (size_t)std::max(mat->n_cores, 1), 0.0);
  for (int c = 0; c < mat->n_cores && mat->cores_xy; ++c) {
    cores[3 * c] = mat->cores_xy[2 * c]; cores[3 * c + 1] = mat->cores_xy[2 * c + 1]; cores[3 * c + 2] = mat->cores_r[c];
  }
  pb->d_cores.upload(ctx, cores);
  const double* epsq = nullptr;
  if (mat->eps_at_quad) { pb->d_epsq.upload(ctx, mat->eps_at_quad, (size_t)6 * pb->dof.T); epsq = pb->d_epsq.p; }
  pb->d_elem.alloc(ctx, (size_t)12 * pb->dof.T);
  launch_element_setup(ctx, pb->d_p.p, pb->d_edofs.p, pb->dof.V, pb->dof.T, *mat, pb->d_cores.p, epsq, pb->d_elem.p);
}

// (k0^2, alpha_p) of the vectorial system or (k0^2, shift of the decoupled Hy copy) of the scalar one
inline double second_coeff(const plfem_material& m) { return m.scalar_mode ? m.scalar_shift : m.alpha_p; }
inline int assembly_mode(const plfem_material& m) { return m.scalar_mode ? 2 : 0; }

void upload_pattern(plfem_ctx* ctx, const Pattern& P, DevPattern& D) {
  D.n = P.n; D.nnz = (int64_t)P.col.size();
  D.rowptr.upload(ctx, P.rowptr); D.col.upload(ctx, P.col); D.old_of_new.upload(ctx, P.old_of_new);
  D.rowidx.alloc(ctx, std::max<size_t>(P.col.size(), 1));
  launch_expand_rows(ctx, D);
}

void ensure_full_pattern(plfem_problem* pb) {
  if (pb->full_ready) return;
  std::vector<int32_t> ident(pb->dof.N);
  for (int64_t i = 0; i < pb->dof.N; ++i) ident[i] = (int32_t)i;
  build_pattern(pb->dof, ident, (int32_t)pb->dof.N, pb->full);
  upload_pattern(pb->ctx, pb->full, pb->dfull);
  pb->full_ready = true;
}

void ensure_adj(plfem_problem* pb) {
  if (pb->adj_ready) return;
  std::vector<int32_t> nid(pb->dof.N, -1);
  for (size_t i = 0; i < pb->dof.interior.size(); ++i) nid[pb->dof.interior[i]] = (int32_t)i;
  build_pattern(pb->dof, nid, (int32_t)pb->dof.interior.size(), pb->adj);
  pb->adj_ready = true;
}

// returns true when a new plan was built
bool ensure_plan(plfem_problem* pb, int leaf_nodes, int max_sn_nodes, bool reuse) {
  SymbolicOptions opt;
  if (leaf_nodes > 0) opt.leaf_nodes = leaf_nodes;
  if (max_sn_nodes > 0) opt.max_sn_nodes = max_sn_nodes;
  need(opt.max_sn_nodes <= 64 && opt.leaf_nodes <= 64, "leaf_nodes and max_sn_nodes must be <= 64");
  if (pb->plan_ready && reuse && opt.leaf_nodes == pb->plan_opt.leaf_nodes && opt.max_sn_nodes == pb->plan_opt.max_sn_nodes) return false;
  const bool need_adj = front_plan_needs_adjacency(pb->dof);
  if (need_adj) ensure_adj(pb);
  const int32_t n = (int32_t)pb->dof.interior.size();
  std::vector<double> x(n), y(n);
  for (int32_t i = 0; i < n; ++i) { x[i] = pb->dof.doflocs[pb->dof.interior[i]]; y[i] = pb->dof.doflocs[pb->dof.N + pb->dof.interior[i]]; }
  build_front_plan(pb->dof, need_adj ? &pb->adj : nullptr, x.data(), y.data(), opt, pb->plan);
  pb->plan_opt = opt;
  pb->plan_ready = true;
  pb->plan_serial++;  // device copies made from the previous plan are stale
  return true;
}

}  // namespace

// ---- C ABI --------------------------------------------------------------------------------------------
extern "C" {

const char* plfem_version(void) { return "plfem 0.1 sm_100a"; }

void plfem_set_host_threads(int n) { plfem::set_host_threads(n); }

int plfem_ctx_set_sweep_schedule(plfem_ctx* ctx, int schedule) {
  if (!ctx || schedule < -1 || schedule > 1) return PLFEM_ERR_INVALID;
  ctx->sweep_schedule = schedule;
  return PLFEM_OK;
}

int plfem_ctx_sweep_schedule(const plfem_ctx* ctx) { return ctx ? (plfem::use_fused_sweeps(ctx) ? 0 : 1) : -1; }

int plfem_ctx_create(int device, plfem_ctx** out) {
  if (!out) return PLFEM_ERR_INVALID;
  *out = nullptr;
  tune_host_allocator_once();
  auto* ctx = new plfem_ctx();
  int st = guarded(ctx, [&] {
    int ndev = 0;
    PLFEM_CUDA(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) throw StatusError(PLFEM_ERR_CUDA, "no such CUDA device: " + std::to_string(device));
    PLFEM_CUDA(cudaSetDevice(device));
    ctx->device = device;
    PLFEM_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    for (auto& e : ctx->ev) PLFEM_CUDA(cudaEventCreate(&e));
    upload_tables();
  });
  if (st != PLFEM_OK) {
    // keep the context alive so the caller can read the message
    static thread_local std::string last;
    last = ctx->err;
    fprintf(stderr, "plfem_ctx_create: %s\n", last.c_str());
    delete ctx;
    return st;
  }
  *out = ctx;
  return PLFEM_OK;
}

void plfem_ctx_destroy(plfem_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  ctx->last_work.reset();        // its buffers go back to the arena before the arena frees everything
  ctx->arena.destroy();
  if (ctx->pinned) cudaFreeHost(ctx->pinned);
  for (auto& e : ctx->ev) if (e) cudaEventDestroy(e);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

const char* plfem_last_error(const plfem_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int plfem_problem_create(plfem_ctx* ctx, const double* p, const int64_t* t, int64_t V, int64_t T, plfem_problem** out) {
  if (!out) return PLFEM_ERR_INVALID;
  *out = nullptr;
  tune_host_allocator_once();
  std::unique_ptr<plfem_problem> pb(new plfem_problem());
  pb->ctx = ctx;
  int st = guarded(ctx, [&] {
    need(p && t, "p or t is NULL");
    need(V >= 3 && T >= 1, "need at least one triangle");
    need(V + 3 * T < (int64_t(1) << 31), "mesh too large for 32-bit DOF ids");
    build_dof_tables(p, t, V, T, pb->dof);
    pb->p_host.assign(p, p + 2 * V);
    {
      uint64_t h = 1469598103934665603ull;     // FNV-1a over the raw bytes of p and t
      auto mix = [&](const void* data, size_t bytes) {
        const uint64_t* w = static_cast<const uint64_t*>(data);
        for (size_t i = 0; i < bytes / 8; ++i) { h ^= w[i]; h *= 1099511628211ull; h ^= h >> 29; }
      };
      mix(p, sizeof(double) * 2 * (size_t)V); mix(t, sizeof(int64_t) * 3 * (size_t)T);
      pb->mesh_print = h;
    }
    if (!ctx) return;   // host-only problem: DOF tables and front plan only (tests of the host logic)
    PLFEM_CUDA(cudaSetDevice(ctx->device));
    pb->d_p.upload(ctx, pb->p_host);
    pb->d_edofs.upload(ctx, pb->dof.edofs);
    pb->d_n2e_ptr.upload(ctx, pb->dof.n2e_ptr);
    pb->d_n2e.upload(ctx, pb->dof.n2e);
  });
  if (st != PLFEM_OK) return st;
  *out = pb.release();
  return PLFEM_OK;
}

int plfem_problem_set_dirichlet(plfem_problem* pb, int on) {
  if (!pb) return PLFEM_ERR_INVALID;
  return guarded(pb->ctx, [&] {
    DofTables& d = pb->dof;
    d.interior.clear();
    if (on) {
      std::vector<uint8_t> is_b(d.N, 0);
      for (int32_t b : d.boundary) is_b[b] = 1;
      for (int32_t i = 0; i < d.N; ++i) if (!is_b[i]) d.interior.push_back(i);
    } else {
      d.interior.resize(d.N);
      for (int32_t i = 0; i < d.N; ++i) d.interior[i] = i;
    }
    pb->plan_ready = false; pb->adj_ready = false; pb->work.reset();     // everything derived from the interior set is stale
  });
}

void plfem_problem_destroy(plfem_problem* pb) {
  if (!pb) return;
  if (pb->ctx) cudaSetDevice(pb->ctx->device);
  delete pb;
}

int plfem_problem_info(const plfem_problem* pb, plfem_mesh_info* info) {
  if (!pb || !info) return PLFEM_ERR_INVALID;
  info->V = pb->dof.V; info->T = pb->dof.T; info->E = pb->dof.E; info->N = pb->dof.N;
  info->n_boundary = (int64_t)pb->dof.boundary.size();
  info->n_interior = (int64_t)pb->dof.interior.size();
  info->nnz_scalar = pb->full_ready ? (int64_t)pb->full.col.size() : -1;
  info->n_degenerate = pb->dof.n_degenerate;
  return PLFEM_OK;
}

int plfem_problem_dofs(const plfem_problem* pb, int64_t* element_dofs, double* doflocs, int64_t* boundary, int64_t* interior) {
  if (!pb) return PLFEM_ERR_INVALID;
  const DofTables& d = pb->dof;
  if (element_dofs)
    for (int64_t e = 0; e < d.T; ++e)
      for (int k = 0; k < 6; ++k) element_dofs[k * d.T + e] = d.edofs[6 * e + k];
  if (doflocs) std::copy(d.doflocs.begin(), d.doflocs.end(), doflocs);
  if (boundary) std::copy(d.boundary.begin(), d.boundary.end(), boundary);
  if (interior) std::copy(d.interior.begin(), d.interior.end(), interior);
  return PLFEM_OK;
}

int plfem_quad_points(const plfem_problem* pb, double* out_xy) {
  if (!pb || !out_xy) return PLFEM_ERR_INVALID;
  const DofTables& d = pb->dof;
  const RefTables& rt = ref_tables();
  const double* px = pb->p_host.data();
  const double* py = px + d.V;
  for (int64_t e = 0; e < d.T; ++e) {
    const int32_t v0 = d.edofs[6 * e], v1 = d.edofs[6 * e + 1], v2 = d.edofs[6 * e + 2];
    const double a00 = px[v1] - px[v0], a01 = px[v2] - px[v0], a10 = py[v1] - py[v0], a11 = py[v2] - py[v0];
    for (int q = 0; q < 6; ++q) {
      volatile double t0 = a00 * rt.qx[q], t1 = a01 * rt.qy[q]; volatile double sx = t0 + t1;
      volatile double t2 = a10 * rt.qx[q], t3 = a11 * rt.qy[q]; volatile double sy = t2 + t3;
      out_xy[e * 6 + q] = sx + px[v0];
      out_xy[(d.T + e) * 6 + q] = sy + py[v0];
    }
  }
  return PLFEM_OK;
}

int plfem_assemble(plfem_problem* pb, const plfem_material* mat) {
  if (!pb || !pb->ctx) return PLFEM_ERR_INVALID;
  plfem_ctx* ctx = pb->ctx;
  return guarded(ctx, [&] {
    PLFEM_CUDA(cudaSetDevice(ctx->device));
    if (pb->dof.n_degenerate > 0)
      throw StatusError(PLFEM_ERR_DEGENERATE, std::to_string(pb->dof.n_degenerate) + " zero-area triangle(s): the affine map is singular");
    ensure_full_pattern(pb);
    need(mat != nullptr, "material is NULL");
    plfem_material vmat = *mat;
    vmat.scalar_mode = 0;          // the export path always assembles the ten scalar matrices of the H-field forms (M among them)
    mat = &vmat;
    upload_material(pb, mat);
    const int64_t nnz = pb->dfull.nnz;
    pb->d_full_vals.alloc(ctx, (size_t)NV_EXPORT * nnz);
    pb->d_full_flags.alloc(ctx, (size_t)nnz);
    launch_assemble(ctx, pb->dfull, pb->d_n2e_ptr.p, pb->d_n2e.p, pb->d_edofs.p, pb->d_elem.p, mat->k0 * mat->k0,
                    mat->alpha_p, 1, pb->d_full_vals.p, pb->d_full_flags.p);
    pb->h_full_vals.resize((size_t)NV_EXPORT * nnz);
    pb->h_full_flags.resize((size_t)nnz);
    pb->d_full_vals.download(pb->h_full_vals.data(), pb->h_full_vals.size());
    pb->d_full_flags.download(pb->h_full_flags.data(), pb->h_full_flags.size());
    PLFEM_CUDA(stream_wait(ctx->stream));
    pb->last_mat = *mat;
    pb->assembled = true;
  });
}

int plfem_export_csr(plfem_problem* pb, int which, int64_t* rows_out, int64_t* nnz_out, int64_t* indptr, int64_t* indices,
                     double* data) {
  if (!pb) return PLFEM_ERR_INVALID;
  plfem_ctx* ctx = pb->ctx;
  return guarded(ctx, [&] {
    if (!pb->assembled) throw StatusError(PLFEM_ERR_NOT_READY, "plfem_export_csr called before plfem_assemble");
    const Pattern& P = pb->full;
    const int64_t N = pb->dof.N, nnz = (int64_t)P.col.size();
    const double* v = pb->h_full_vals.data();
    const uint32_t* fl = pb->h_full_flags.data();
    const double k0sq = pb->last_mat.k0 * pb->last_mat.k0, al = pb->last_mat.alpha_p;
    auto val = [&](int k, int64_t z) { return v[(int64_t)k * nnz + z]; };
    // scalar value + keep rule for block (bi, bj) of A, or for a scalar matrix
    auto a_block = [&](int bi, int bj, int64_t z, double& out) {
      volatile double t;
      if (bi == 0 && bj == 0) { t = val(X_KXX, z) + al * val(X_DXX, z); t = t - k0sq * val(X_M, z); }
      else if (bi == 1 && bj == 1) { t = val(X_KYY, z) + al * val(X_DYY, z); t = t - k0sq * val(X_M, z); }
      else if (bi == 0) t = val(X_KXY, z) + al * val(X_DXY, z);
      else t = val(X_KYX, z) + al * val(X_DYX, z);
      out = t;
      return out != 0.0;   // sparse +/- drop exact zeros, NaN is kept
    };
    const bool interior_only = (which == PLFEM_MAT_A_INT || which == PLFEM_MAT_B_INT);
    const bool is_a = (which == PLFEM_MAT_A || which == PLFEM_MAT_A_INT);
    const bool is_b = (which == PLFEM_MAT_B || which == PLFEM_MAT_B_INT);
    std::vector<int32_t> imap;  // scalar node -> position among kept nodes
    int64_t nk = N;
    if (interior_only) {
      imap.assign(N, -1);
      for (size_t i = 0; i < pb->dof.interior.size(); ++i) imap[pb->dof.interior[i]] = (int32_t)i;
      nk = (int64_t)pb->dof.interior.size();
    }
    auto kept = [&](int64_t node) { return interior_only ? imap[node] : (int32_t)node; };
    int scalar_k = -1;
    switch (which) {
      case PLFEM_MAT_DXX: scalar_k = X_DXX; break; case PLFEM_MAT_DYY: scalar_k = X_DYY; break;
      case PLFEM_MAT_DXY: scalar_k = X_DXY; break; case PLFEM_MAT_MINV: scalar_k = X_MINV; break;
      case PLFEM_MAT_KXX: scalar_k = X_KXX; break; case PLFEM_MAT_KYY: scalar_k = X_KYY; break;
      case PLFEM_MAT_KXY: scalar_k = X_KXY; break; case PLFEM_MAT_KYX: scalar_k = X_KYX; break;
      case PLFEM_MAT_M: scalar_k = X_M; break;
      default: break;
    }
    if (!is_a && !is_b && scalar_k < 0) throw StatusError(PLFEM_ERR_INVALID, "unknown matrix id");
    const int nblk = (is_a || is_b) ? 2 : 1;
    const int64_t rows = nblk * nk;
    if (rows_out) *rows_out = rows;
    int64_t count = 0;
    const bool fill = (indptr && indices && data);
    if (fill) indptr[0] = 0;
    for (int bi = 0; bi < nblk; ++bi)
      for (int64_t r = 0; r < N; ++r) {
        const int32_t kr = kept(r);
        if (kr < 0) continue;
        for (int bj = 0; bj < nblk; ++bj) {
          if (is_b && bi != bj) continue;
          for (int32_t z = P.rowptr[r]; z < P.rowptr[r + 1]; ++z) {
            const int32_t kc = kept(P.col[z]);
            if (kc < 0) continue;
            double x;
            bool keep;
            if (is_a) keep = a_block(bi, bj, z, x);
            else { const int k = is_b ? X_MINV : scalar_k; x = val(k, z); keep = (fl[z] >> k) & 1u; }
            if (!keep) continue;
            if (fill) { indices[count] = bj * nk + kc; data[count] = x; }
            ++count;
          }
        }
        if (fill) indptr[bi * nk + kr + 1] = count;
      }
    if (nnz_out) *nnz_out = count;
  });
}

int plfem_spmv_csr(plfem_ctx* ctx, int64_t rows, int64_t nnz, const int64_t* indptr, const int64_t* indices,
                   const double* data, const double* x, double* y, int repeat, float* ms) {
  if (!ctx) return PLFEM_ERR_INVALID;
  return guarded(ctx, [&] {
    need(indptr && indices && data && x && y, "NULL array");
    need(rows > 0 && nnz >= 0 && nnz < (int64_t(1) << 31), "bad sizes");
    PLFEM_CUDA(cudaSetDevice(ctx->device));
    std::vector<int32_t> rp(rows + 1), ci(nnz);
    int64_t maxcol = rows;
    for (int64_t i = 0; i <= rows; ++i) rp[i] = (int32_t)indptr[i];
    for (int64_t i = 0; i < nnz; ++i) { need(indices[i] >= 0, "negative column"); ci[i] = (int32_t)indices[i]; maxcol = std::max(maxcol, indices[i] + 1); }
    need(maxcol == rows, "plfem_spmv_csr expects a square matrix");
    DevBuf<int32_t> d_rp, d_ci; DevBuf<double> d_v, d_x, d_y;
    d_rp.upload(ctx, rp); d_ci.upload(ctx, ci); d_v.upload(ctx, data, nnz); d_x.upload(ctx, x, rows); d_y.alloc(ctx, rows);
    const int reps = std::max(repeat, 1);
    launch_spmv_csr(ctx, rows, d_rp.p, d_ci.p, d_v.p, d_x.p, d_y.p);  // warm-up
    PLFEM_CUDA(cudaEventRecord(ctx->ev[0], ctx->stream));
    for (int i = 0; i < reps; ++i) launch_spmv_csr(ctx, rows, d_rp.p, d_ci.p, d_v.p, d_x.p, d_y.p);
    PLFEM_CUDA(cudaEventRecord(ctx->ev[1], ctx->stream));
    d_y.download(y, rows);
    PLFEM_CUDA(stream_wait(ctx->stream));
    float t = 0;
    PLFEM_CUDA(cudaEventElapsedTime(&t, ctx->ev[0], ctx->ev[1]));
    if (ms) *ms = t / reps;
  });
}

// Modal solves of nb designs as ONE forest: per-design symbolic analysis on host threads, then a single
// chain of launches (assembly slices, one level-batched factorisation, lockstep block Lanczos, reductions).
// A failure of one design (singular shift, no convergence) is reported in statuses[b] and leaves the others intact.
static void solve_forest(plfem_ctx* ctx, int nb, plfem_problem* const* pbs, const plfem_material* mats, const plfem_solve_opts* opts,
                         double* const* eigvals, double* const* evecs, double* const* metrics, int32_t* core_counts,
                         plfem_solve_stats* stats, int32_t* statuses, std::vector<std::string>& errs, SolveWork& W, bool reuse_work) {
  PLFEM_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  std::vector<int64_t> nint(nb);
  int kmax = 0;
  for (int b = 0; b < nb; ++b) {
    plfem_problem* pb = pbs[b];
    need(pb && pb->ctx == ctx, "every problem of a batch must belong to the batch's context");
    need(eigvals[b] && metrics[b], "NULL output");
    if (pb->dof.n_degenerate > 0)
      throw StatusError(PLFEM_ERR_DEGENERATE, std::to_string(pb->dof.n_degenerate) + " zero-area triangle(s): the affine map is singular");
    nint[b] = (int64_t)pb->dof.interior.size();
    need(opts[b].k >= 1 && opts[b].k < 2 * nint[b], "k out of range");
    for (int c = 0; c < b; ++c) need(pbs[c] != pb, "a problem appears twice in the batch");
    kmax = std::max(kmax, opts[b].k);
  }
  const plfem_solve_opts& o0 = opts[0];
  const int refine = o0.refine == 0 ? 1 : std::max(o0.refine, 0);     // single-vector path
  const int refine_block = o0.refine == 0 ? -1 : std::max(o0.refine, 0);   // block path: 0 in the options = choose by probing
  const int block = o0.block == 0 ? SOLVE_NRHS : o0.block;     // 0 = default (block Lanczos), 1 = single vector
  need(block == 1 || block == SOLVE_NRHS, "block must be 0 (default), 1 or SOLVE_NRHS");
  const bool single_vector = (nb == 1) && (block == 1 || 2 * nint[0] < 8 * SOLVE_NRHS);
  need(nb == 1 || block == SOLVE_NRHS, "a batch of designs needs the block eigensolver");
  ctx->launches = 0;
  const double t0 = now_ms();
  static const bool cpu_timing = std::getenv("PLFEM_TIMING") != nullptr;
  double c_[6] = {cpu_ms(), 0, 0, 0, 0, 0};

  // -- symbolic (host), designs in parallel -------------------------------------------------------------
  std::vector<double> ms_sym(nb, 0.0);
  std::vector<std::vector<uint8_t>> masks(nb);
  const bool have_work = reuse_work && W.ready;
  // Designs that allow it (reuse_symbolic = 1) and sit on the SAME mesh — the bands of a wavelength sweep, README.md:226-243:
  // the mesh depends on the geometry only, mesh.py:232-289 — share one analysis: the first of them is analysed, the others
  // copy its front plan (ordering, fronts, update sets: integers only, nothing numerical is shared).
  std::vector<int> donor(nb, -1);
  for (int b = 0; b < nb; ++b) {
    if (!opts[b].reuse_symbolic || pbs[b]->plan_ready) continue;
    for (int c = 0; c < b; ++c)
      if (opts[c].reuse_symbolic && donor[c] < 0 && pbs[c]->mesh_print == pbs[b]->mesh_print && pbs[c]->dof.N == pbs[b]->dof.N &&
          pbs[c]->dof.T == pbs[b]->dof.T && pbs[c]->dof.interior.size() == pbs[b]->dof.interior.size() &&
          opts[c].leaf_nodes == opts[b].leaf_nodes && opts[c].max_sn_nodes == opts[b].max_sn_nodes) { donor[b] = c; break; }
  }
  {
    const int hw = host_thread_budget();      // threads this call may occupy (lowered when several forests / ranks share the host)
    const int outer = std::min(nb, hw);
    std::atomic<int> next{0};
    std::exception_ptr err; std::mutex mu;
    auto run_parallel = [&](auto&& body) {
      next = 0;
      auto work = [&] {
        if (nb > 1) set_host_threads_local(std::max(1, std::min(8, hw / outer)));
        for (int b; (b = next.fetch_add(1)) < nb;) {
          try { body(b); } catch (...) { std::lock_guard<std::mutex> g(mu); if (!err) err = std::current_exception(); }
        }
        set_host_threads_local(0);
      };
      std::vector<std::thread> th;
      for (int t = 1; t < outer; ++t) th.emplace_back(work);
      work();
      for (auto& t : th) t.join();
      if (err) std::rethrow_exception(err);
    };
    run_parallel([&](int b) {            // ordering + front plan of every design that does not borrow one
      if (donor[b] >= 0) return;
      const double ta = now_ms();
      ensure_plan(pbs[b], opts[b].leaf_nodes, opts[b].max_sn_nodes, opts[b].reuse_symbolic != 0);
      ms_sym[b] = now_ms() - ta;
    });
    for (int b = 0; b < nb; ++b) {
      if (donor[b] < 0) continue;
      plfem_problem* pb = pbs[b];
      pb->plan = pbs[donor[b]]->plan; pb->plan_opt = pbs[donor[b]]->plan_opt;
      pb->plan_ready = true; pb->plan_serial++;
    }
    run_parallel([&](int b) {            // core mask of the permuted interior nodes (solver_fem.py:200-203)
      const double ta = now_ms();
      plfem_problem* pb = pbs[b];
      const plfem_material* mat = &mats[b];
      const int64_t n = nint[b];
      std::vector<uint8_t>& mask = masks[b];
      mask.assign(n, 0);
      const double* Xc = pb->dof.doflocs.data();
      const double* Yc = Xc + pb->dof.N;
      int32_t cnt = 0;
      for (int64_t r = 0; r < n; ++r) {
        const int32_t node = pb->dof.interior[pb->plan.perm[r]];
        bool in = false;
        for (int c = 0; c < mat->n_cores && mat->cores_xy; ++c) {
          // (outside the core's bounding square, widened by 1e-9 relative so that no rounding of the test below can matter:
          //  |dx| > r (1 + 1e-9) implies fl(dx*dx) > fl(r*r), and fl(dx2 + dy2) >= dx2)
          const double rb = std::fabs(mat->cores_r[c]) * (1.0 + 1e-9);
          if (std::fabs(Xc[node] - mat->cores_xy[2 * c]) > rb || std::fabs(Yc[node] - mat->cores_xy[2 * c + 1]) > rb) continue;
          volatile double dx = Xc[node] - mat->cores_xy[2 * c], dy = Yc[node] - mat->cores_xy[2 * c + 1];
          volatile double dx2 = dx * dx, dy2 = dy * dy, rr = mat->cores_r[c] * mat->cores_r[c];
          volatile double d2 = dx2 + dy2;
          if (d2 <= rr) in = true;
        }
        mask[r] = in; cnt += in;
      }
      if (core_counts) core_counts[b] = cnt;
      ms_sym[b] += now_ms() - ta;
    });
  }
  c_[1] = cpu_ms();
  if (!have_work) {
    std::vector<const FrontPlan*> plans(nb);
    for (int b = 0; b < nb; ++b) plans[b] = &pbs[b]->plan;
    std::vector<int32_t> node_off;
    static const bool timing = std::getenv("PLFEM_TIMING") != nullptr;
    const double ta = now_ms();
    merge_front_plans(plans, W.plan, node_off, W.front_off);
    const double tb = now_ms();
    W.bd.set(ctx, node_off);
    {
      // sparsity pattern in elimination order: built on the device from the resident element tables
      std::vector<PatternSource> src(nb);
      std::vector<int32_t> old_of_new((size_t)node_off[nb]);
      for (int b = 0; b < nb; ++b) {
        const plfem_problem* pb = pbs[b];
        src[b] = {pb->d_n2e_ptr.p, pb->d_n2e.p, pb->d_edofs.p, pb->dof.N};
        int32_t* o = old_of_new.data() + node_off[b];
        for (int32_t r = 0; r < pb->plan.n; ++r) o[r] = pb->dof.interior[pb->plan.perm[r]];
      }
      build_device_pattern(ctx, nb, src.data(), node_off, old_of_new, W.dpat, W.nnz_off);
    }
    const double tc = now_ms();
    build_dev_plan(ctx, W.plan, W.dplan);
    W.d_perm.upload(ctx, W.plan.perm);
    W.ready = true;
    if (timing) fprintf(stderr, "[plfem] forest of %d: merge %.2f ms, device pattern %.2f ms, device plan %.2f ms\n", nb, tb - ta, tc - tb, now_ms() - tc);
  }
  const BatchDims& bd = W.bd;
  const int64_t n_tot = bd.noff[nb];
  const double t1 = now_ms();
  c_[2] = cpu_ms();

  // -- assembly: one slice of the concatenated value arrays per design ----------------------------------------
  PLFEM_CUDA(cudaEventRecord(ctx->ev[0], st));
  const int64_t nnz = W.dpat.nnz;
  W.d_vals.alloc(ctx, (size_t)NV_SOLVE * nnz);
  std::vector<double> sig(n_tot);
  std::vector<uint8_t> mask_all(n_tot);
  W.asm_args.resize(nb);
  for (int b = 0; b < nb; ++b) {
    plfem_problem* pb = pbs[b];
    upload_material(pb, &mats[b]);
    W.asm_args[b] = {pb->d_p.p, pb->d_edofs.p, pb->d_n2e_ptr.p, pb->d_n2e.p, pb->d_elem.p, pb->d_cores.p, pb->dof.V, pb->dof.T, mats[b]};
    W.asm_args[b].mat.cores_xy = nullptr; W.asm_args[b].mat.cores_r = nullptr; W.asm_args[b].mat.eps_at_quad = nullptr;
    launch_assemble_slice(ctx, bd.noff[b], bd.noff[b + 1] - bd.noff[b], W.dpat, pb->d_n2e_ptr.p, pb->d_n2e.p, pb->d_edofs.p, pb->d_elem.p,
                          mats[b].k0 * mats[b].k0, second_coeff(mats[b]), assembly_mode(mats[b]), W.d_vals.p, nnz, nullptr);
    std::fill(sig.begin() + bd.noff[b], sig.begin() + bd.noff[b + 1], opts[b].sigma);
    std::copy(masks[b].begin(), masks[b].end(), mask_all.begin() + bd.noff[b]);
  }
  W.d_sigma.upload(ctx, sig);
  W.d_in_core.upload(ctx, mask_all);
  PLFEM_CUDA(cudaEventRecord(ctx->ev[1], st));

  // -- numeric factorisation of A - sigma B, all designs level by level ----------------------------------------
  launch_front_load(ctx, W.dpat, W.dplan, W.d_vals.p, W.d_sigma.p);
  run_factorization(ctx, W.dplan);
  PLFEM_CUDA(cudaEventRecord(ctx->ev[2], st));
  int32_t fstat[4] = {0, 0, 0, 0};
  W.dplan.status.download(fstat, 4);
  PLFEM_CUDA(stream_wait(st));   // also covers the pageable sig / mask uploads above
  std::vector<DesignEig> des(nb);
  for (int b = 0; b < nb; ++b) { des[b].k = opts[b].k; des[b].sigma = opts[b].sigma; des[b].tol = opts[b].tol > 0 ? opts[b].tol : 1e-7; }
  if (fstat[0]) {
    const int bad = (int)(std::upper_bound(W.front_off.begin(), W.front_off.end(), fstat[3]) - W.front_off.begin()) - 1;
    const int bb = std::max(0, std::min(nb - 1, bad));
    des[bb].status = PLFEM_ERR_SINGULAR; des[bb].err = "a pivot block of A - sigma*B is exactly singular or non-finite";
  }

  // -- eigensolver ----------------------------------------------------------------------------------
  c_[3] = cpu_ms();
  DevBuf<double> d_v0;
  const double* v0p = nullptr;
  bool any_v0 = false;
  for (int b = 0; b < nb; ++b) any_v0 |= (opts[b].v0 != nullptr);
  if (any_v0) {
    std::vector<double> v0(2 * n_tot, 1.0);
    for (int b = 0; b < nb; ++b) {
      if (!opts[b].v0) continue;
      const int64_t n = nint[b], r0 = bd.noff[b];
      for (int64_t r = 0; r < n; ++r) { const int32_t ip = pbs[b]->plan.perm[r]; v0[2 * (r0 + r)] = opts[b].v0[ip]; v0[2 * (r0 + r) + 1] = opts[b].v0[n + ip]; }
    }
    d_v0.upload(ctx, v0);
    PLFEM_CUDA(stream_wait(st));
    v0p = d_v0.p;
  }
  DevBuf<double> X;
  int64_t ldx = 2 * n_tot;
  EigenResult er;
  int maxiter = 0, ncv_req = 0;
  for (int b = 0; b < nb; ++b) { maxiter = std::max(maxiter, opts[b].maxiter > 0 ? opts[b].maxiter : 12000); ncv_req = std::max(ncv_req, opts[b].ncv); }
  if (des[0].status != PLFEM_OK && nb == 1) {
    // nothing to iterate on
  } else if (!single_vector) {
    run_eigensolver_block(ctx, W.dpat, W.dplan, W.d_vals.p, W.d_sigma.p, bd, des, ncv_req > 0 ? ncv_req : 3 * kmax, maxiter, refine_block, v0p, X, er);
  } else {
    const int64_t n = nint[0];
    const int k = opts[0].k;
    int ncv = o0.ncv > 0 ? o0.ncv : std::max(2 * k + 1, 20);
    ncv = (int)std::min<int64_t>(ncv, 2 * n);
    need(ncv > k + 1, "ncv must exceed k + 1");
    try {
      run_eigensolver(ctx, W.dpat, W.dplan, W.d_vals.p, W.d_sigma.p, o0.sigma, k, ncv, des[0].tol, maxiter, refine, v0p, X, des[0].lambda, er);
      des[0].nconv = er.nconv;
    } catch (const StatusError& e) {
      if (e.status != PLFEM_ERR_NO_CONVERGENCE && e.status != PLFEM_ERR_SINGULAR) throw;
      des[0].status = e.status; des[0].err = e.what(); des[0].nconv = er.nconv;
    }
    des[0].n_block_op = 0; des[0].n_restart = er.n_restart;
  }
  c_[4] = cpu_ms();
  PLFEM_CUDA(cudaEventRecord(ctx->ev[3], st));
  W.dplan.status.download(fstat, 4);
  PLFEM_CUDA(stream_wait(st));
  if (fstat[1]) throw StatusError(PLFEM_ERR_INTERNAL, "a sweep kernel gave up waiting for a bulk copy of its factor stream");

  // -- per-mode reductions + eigenvectors in reference ordering, design by design ---------------------------
  std::vector<std::vector<double>> resid(nb);
  std::vector<DevBuf<double>> d_ev(nb), d_met(nb), d_res(nb);
  for (int b = 0; b < nb; ++b) {
    const int k = des[b].k;
    if (des[b].lambda.size() != (size_t)k) continue;     // failed before any Ritz pair existed
    const int64_t n = nint[b];
    if (evecs[b]) d_ev[b].alloc(ctx, (size_t)k * 2 * n);
    d_met[b].alloc(ctx, (size_t)k * PLFEM_NMETRICS); d_res[b].alloc(ctx, (size_t)2 * k);
    run_mode_metrics(ctx, W.dpat, W.d_vals.p, W.d_perm.p, W.d_in_core.p, X.p, ldx, bd.noff[b], (int32_t)n, des[b].lambda, k,
                     evecs[b] ? d_ev[b].p : nullptr, d_met[b].p, d_res[b].p);
    resid[b].resize(2 * k);
    d_met[b].download(metrics[b], (size_t)k * PLFEM_NMETRICS);
    d_res[b].download(resid[b].data(), resid[b].size());
    if (evecs[b]) d_ev[b].download(evecs[b], (size_t)k * 2 * n);
  }
  PLFEM_CUDA(cudaEventRecord(ctx->ev[4], st));
  PLFEM_CUDA(stream_wait(st));
  const double t2 = now_ms();
  c_[5] = cpu_ms();
  if (cpu_timing)
    fprintf(stderr, "[plfem] process CPU ms of a forest of %d: analysis %.1f, merge + device plans %.1f, assembly + factorisation issue %.1f, "
                    "eigensolver (launches + convergence checks) %.1f, reductions + copies %.1f; total %.1f = %.2f per design\n",
            nb, c_[1] - c_[0], c_[2] - c_[1], c_[3] - c_[2], c_[4] - c_[3], c_[5] - c_[4], c_[5] - c_[0], (c_[5] - c_[0]) / nb);

  float ms_asm = 0, ms_fac = 0, ms_lan = 0, ms_met = 0;
  cudaEventElapsedTime(&ms_asm, ctx->ev[0], ctx->ev[1]);
  cudaEventElapsedTime(&ms_fac, ctx->ev[1], ctx->ev[2]);
  cudaEventElapsedTime(&ms_lan, ctx->ev[2], ctx->ev[3]);
  cudaEventElapsedTime(&ms_met, ctx->ev[3], ctx->ev[4]);
  for (int b = 0; b < nb; ++b) {
    // The Lanczos convergence test assumes an exact operator; the block-LDL^T solve (+ the refinement steps the probe chose)
    // is not.  The TRUE normwise backward error ||A x - lambda B x|| / ((||A||_F + |lambda| ||B||_F) ||x||) of every returned
    // pair is therefore checked against the tolerance before a design is reported as PLFEM_OK.
    double mr = 0.0;
    for (size_t i = 0; i + 1 < resid[b].size(); i += 2) {
      const double r = resid[b][i] / (resid[b][i + 1] + 1e-300);
      mr = (r > mr || !(r == r)) ? r : mr;          // a NaN residual must not pass
    }
    if (des[b].status == PLFEM_OK && !resid[b].empty() && !(mr <= 0.1 * des[b].tol)) {
      des[b].status = PLFEM_ERR_NO_CONVERGENCE;
      des[b].err = "returned eigenpairs miss the backward-error bar: max ||A x - lambda B x|| / ((||A|| + |lambda| ||B||) ||x||) = " +
                   std::to_string(mr) + " > 0.1 * tol (the factorisation of A - sigma*B is less accurate than the refinement probe estimated)";
    }
    statuses[b] = des[b].status;
    errs[b] = des[b].err;
    if (des[b].lambda.size() == (size_t)des[b].k) std::copy(des[b].lambda.begin(), des[b].lambda.end(), eigvals[b]);
    if (!stats) continue;
    plfem_solve_stats* s = &stats[b];
    std::memset(s, 0, sizeof(*s));
    const FrontPlan& P = pbs[b]->plan;
    s->nconv = des[b].nconv; s->n_op = single_vector ? er.n_op : des[b].n_block_op * SOLVE_NRHS;
    s->n_restart = des[b].n_restart; s->n_block_op = des[b].n_block_op;
    s->n_fronts = P.nfronts; s->n_levels = P.nlevels; s->max_front_nodes = P.max_front;
    s->factor_entries = P.factor_entries; s->front_pool_doubles = P.foff[P.nfronts]; s->factor_flops = P.factor_flops;
    s->max_residual = mr;
    s->ms_symbolic = (float)ms_sym[b];
    s->ms_assemble = ms_asm; s->ms_factor = ms_fac; s->ms_lanczos = ms_lan; s->ms_metrics = ms_met;
    s->ms_total = (float)(t2 - t0);
    s->kernel_launches = ctx->launches;
    s->batch_size = nb; s->batch_block_ops = er.n_block_op; s->ms_symbolic_wall = (float)(t1 - t0);
    s->refine_steps = single_vector ? refine : er.refine_steps;
    s->probe_rho = des[b].solve_residual;
  }
}

int plfem_solve_modes(plfem_problem* pb, const plfem_material* mat, const plfem_solve_opts* o, double* eigvals, double* evecs,
                      double* metrics, int32_t* core_dof_count, plfem_solve_stats* stats) {
  if (!pb || !pb->ctx) return PLFEM_ERR_INVALID;
  plfem_ctx* ctx = pb->ctx;
  return guarded(ctx, [&] {
    need(mat && o && eigvals && metrics, "NULL argument");
    SymbolicOptions want;
    if (o->leaf_nodes > 0) want.leaf_nodes = o->leaf_nodes;
    if (o->max_sn_nodes > 0) want.max_sn_nodes = o->max_sn_nodes;
    const bool reuse = o->reuse_symbolic != 0 && pb->work && pb->work->ready && pb->plan_ready &&
                       pb->work->plan_serial == pb->plan_serial && want.leaf_nodes == pb->plan_opt.leaf_nodes &&
                       want.max_sn_nodes == pb->plan_opt.max_sn_nodes;
    if (!reuse) pb->work = std::make_shared<SolveWork>();
    ctx->last_work = pb->work;
    int32_t status = 0;
    std::vector<std::string> errs(1);
    plfem_problem* pbs[1] = {pb};
    double* ev[1] = {eigvals}; double* vc[1] = {evecs}; double* mt[1] = {metrics};
    solve_forest(ctx, 1, pbs, mat, o, ev, vc, mt, core_dof_count, stats, &status, errs, *pb->work, reuse);
    pb->work->plan_serial = pb->plan_serial;
    if (status != PLFEM_OK) throw StatusError(status, errs[0]);
  });
}

int plfem_solve_modes_batch(plfem_ctx* ctx, int32_t nb, plfem_problem* const* pbs, const plfem_material* mats,
                            const plfem_solve_opts* opts, double* const* eigvals, double* const* evecs, double* const* metrics,
                            int32_t* core_dof_counts, plfem_solve_stats* stats, int32_t* statuses) {
  if (!ctx) return PLFEM_ERR_INVALID;
  return guarded(ctx, [&] {
    need(nb >= 1 && pbs && mats && opts && eigvals && evecs && metrics && statuses, "NULL argument");
    auto Wp = std::make_shared<SolveWork>();
    ctx->last_work = Wp;         // kept for plfem_profile_last; replaced (and its buffers recycled) by the next solve
    SolveWork& W = *Wp;
    std::vector<std::string> errs(nb);
    for (int b = 0; b < nb; ++b) statuses[b] = PLFEM_ERR_INTERNAL;
    solve_forest(ctx, nb, pbs, mats, opts, eigvals, evecs, metrics, core_dof_counts, stats, statuses, errs, W, false);
    ctx->err.clear();
    for (int b = 0; b < nb; ++b)
      if (statuses[b] != PLFEM_OK) { ctx->err = "design " + std::to_string(b) + ": " + errs[b]; break; }
  });
}

// Per-kernel timings for the roofline report: CUDA events on the library's own stream, L2 flushed
// (a 256 MiB memset) before every timed repetition, on the device state the last solve of this context left
// behind (a single design or a forest: every launch then carries all its designs).
// out_ms[0] element setup + assembly, [1] front load + factorisation, [2] forward sweep (all levels),
// [3] backward sweep (all levels), [4] B product (spmm), [5] K residual (refinement spmv), [6] / [7] the sweeps
// with SOLVE_NRHS right-hand sides; out_bytes holds the algorithmic bytes of the same items.
static void profile_work(plfem_ctx* ctx, SolveWork& W, int repeat, double* out_ms, double* out_bytes) {
  if (!W.ready || W.d_vals.p == nullptr || W.asm_args.empty()) throw StatusError(PLFEM_ERR_NOT_READY, "run a modal solve first");
  PLFEM_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const int reps = std::max(repeat, 1);
  const int nb = W.bd.nb;
  const int64_t n = W.dpat.n, nnz = W.dpat.nnz, m = 2 * n;
  DevBuf<double> flush, b, x, t;
  flush.alloc(ctx, (size_t)32 << 20);   // 256 MiB > 126 MB L2
  b.alloc(ctx, m); x.alloc(ctx, m); t.alloc(ctx, m);
  std::vector<double> ones(m, 1.0);
  b.upload(ctx, ones);
  auto timed = [&](auto&& body) {
    double tot = 0.0;
    for (int r = 0; r < reps; ++r) {
      PLFEM_CUDA(cudaMemsetAsync(flush.p, r & 0xff, flush.n * sizeof(double), st));
      PLFEM_CUDA(cudaEventRecord(ctx->ev[5], st));
      body();
      PLFEM_CUDA(cudaEventRecord(ctx->ev[6], st));
      PLFEM_CUDA(stream_wait(st));
      float ms = 0; PLFEM_CUDA(cudaEventElapsedTime(&ms, ctx->ev[5], ctx->ev[6]));
      tot += ms;
    }
    return tot / reps;
  };
  out_ms[0] = timed([&] {
    for (int d = 0; d < nb; ++d) {
      const SolveWork::AsmArgs& a = W.asm_args[d];
      launch_element_setup(ctx, a.d_p, a.d_edofs, a.V, a.T, a.mat, a.d_cores, nullptr, a.d_elem);
      launch_assemble_slice(ctx, W.bd.noff[d], W.bd.noff[d + 1] - W.bd.noff[d], W.dpat, a.d_n2e_ptr, a.d_n2e, a.d_edofs, a.d_elem,
                            a.mat.k0 * a.mat.k0, second_coeff(a.mat), assembly_mode(a.mat), W.d_vals.p, nnz, nullptr);
    }
  });
  // mesh in (coordinates + 6 DOF ids per element), every assembled value out once
  out_bytes[0] = 8.0 * NV_SOLVE * nnz;
  for (const SolveWork::AsmArgs& a : W.asm_args) out_bytes[0] += 16.0 * a.V + 24.0 * a.T;
  out_ms[1] = timed([&] { launch_front_load(ctx, W.dpat, W.dplan, W.d_vals.p, W.d_sigma.p); run_factorization(ctx, W.dplan); });
  out_bytes[1] = 8.0 * 5 * nnz + 8.0 * W.plan.factor_entries;   // read A,B values, write the factor
  out_ms[2] = timed([&] { run_solve_forward(ctx, W.dplan, b.p, x.p); });
  out_bytes[2] = 8.0 * W.plan.factor_entries + 16.0 * m;
  out_ms[3] = timed([&] { run_solve_backward(ctx, W.dplan, x.p); });
  {
    double w = 0; for (int f = 0; f < W.plan.nfronts; ++f) w += 4.0 * W.plan.s[f] * (W.plan.sptr[f + 1] - W.plan.sptr[f]);
    out_bytes[3] = 8.0 * w + 16.0 * m;
  }
  out_ms[4] = timed([&] { launch_spmm_b(ctx, W.dpat, W.d_vals.p, b.p, t.p); });
  out_bytes[4] = 12.0 * nnz + 4.0 * (n + 1) + 32.0 * n;           // values + columns + row pointers + x and y (2 comps)
  out_ms[5] = timed([&] { launch_resid_k(ctx, W.dpat, W.d_vals.p, W.d_sigma.p, x.p, b.p, t.p); });
  out_bytes[5] = (5 * 8.0 + 4.0) * nnz + 4.0 * (n + 1) + 48.0 * n;
  {
    DevBuf<double> b4, x4;
    b4.alloc(ctx, (size_t)m * SOLVE_NRHS); x4.alloc(ctx, (size_t)m * SOLVE_NRHS);
    for (int r = 0; r < SOLVE_NRHS; ++r) PLFEM_CUDA(cudaMemcpyAsync(b4.p + r * m, b.p, m * sizeof(double), cudaMemcpyDeviceToDevice, st));
    out_ms[6] = timed([&] { run_solve_forward(ctx, W.dplan, b4.p, x4.p, SOLVE_NRHS); });
    out_bytes[6] = out_bytes[2] + 16.0 * m * (SOLVE_NRHS - 1);
    out_ms[7] = timed([&] { run_solve_backward(ctx, W.dplan, x4.p, SOLVE_NRHS); });
    out_bytes[7] = out_bytes[3] + 16.0 * m * (SOLVE_NRHS - 1);
    if (const char* tf = std::getenv("PLFEM_TRACE_FILE")) {
      // stage clock of one dataflow forward sweep (file `tf`) and of the backward sweep that follows it (`tf`.bwd):
      // [int64 ntasks, nlevels][int32 level ranges[nlevels + 1]][int64 stamps[ntasks][8]]
      for (int dir = 0; dir < 2; ++dir) {
        const StreamPlan& SP = W.dplan.st;
        const int64_t nt = (int64_t)(dir == 0 ? SP.ftasks.n : SP.btasks.n), nl = W.dplan.nlevels;
        DevBuf<long long> tr;
        tr.alloc(ctx, (size_t)std::max<int64_t>(nt, 1) * 8); tr.zero();
        PLFEM_CUDA(stream_wait(st));
        PLFEM_CUDA(cudaMemsetAsync(flush.p, 1, flush.n * sizeof(double), st));
        if (dir == 0) {
          set_sweep_trace(tr.p);
          run_solve_forward(ctx, W.dplan, b4.p, x4.p, SOLVE_NRHS);
        } else {
          run_solve_forward(ctx, W.dplan, b4.p, x4.p, SOLVE_NRHS);
          PLFEM_CUDA(stream_wait(st));
          PLFEM_CUDA(cudaMemsetAsync(flush.p, 2, flush.n * sizeof(double), st));
          set_sweep_trace(tr.p);
          run_solve_backward(ctx, W.dplan, x4.p, SOLVE_NRHS);
        }
        PLFEM_CUDA(stream_wait(st));
        set_sweep_trace(nullptr);
        std::vector<long long> h((size_t)nt * 8);
        tr.download(h.data(), h.size());
        PLFEM_CUDA(stream_wait(st));
        // level ranges in task order: forward bottom level first; backward top level first (stored reversed: see build_level_plan)
        std::vector<int32_t> ranges(nl + 1);
        for (int l = 0; l <= nl; ++l) ranges[l] = dir == 0 ? SP.fptr[l] : SP.bptr[nl - l];
        const std::string name = std::string(tf) + (dir == 0 ? "" : ".bwd");
        if (FILE* fh = std::fopen(name.c_str(), "wb")) {
          std::fwrite(&nt, 8, 1, fh); std::fwrite(&nl, 8, 1, fh);
          std::fwrite(ranges.data(), 4, (size_t)nl + 1, fh);
          std::fwrite(h.data(), 8, h.size(), fh);
          std::fclose(fh);
        }
      }
    }
  }
}

int plfem_profile_kernels(plfem_problem* pb, const plfem_material* mat, double sigma, int repeat, double* out_ms,
                          double* out_bytes) {
  if (!pb || !pb->ctx) return PLFEM_ERR_INVALID;
  plfem_ctx* ctx = pb->ctx;
  return guarded(ctx, [&] {
    need(mat && out_ms && out_bytes, "NULL argument");
    (void)sigma;   // material and shift of the last solve are still on the device
    if (!pb->work) throw StatusError(PLFEM_ERR_NOT_READY, "run plfem_solve_modes first");
    profile_work(ctx, *pb->work, repeat, out_ms, out_bytes);
  });
}

int plfem_profile_last(plfem_ctx* ctx, int repeat, double* out_ms, double* out_bytes, int32_t* batch_size) {
  if (!ctx) return PLFEM_ERR_INVALID;
  return guarded(ctx, [&] {
    need(out_ms && out_bytes, "NULL argument");
    if (!ctx->last_work) throw StatusError(PLFEM_ERR_NOT_READY, "no solve has run on this context");
    profile_work(ctx, *ctx->last_work, repeat, out_ms, out_bytes);
    if (batch_size) *batch_size = ctx->last_work->bd.nb;
  });
}

// Test hook: x = (A - sigma B)^-1 b with the factors left on the device by the last plfem_solve_modes
// (same sigma), b and x in reference ordering, `refine` refinement steps.  Lets the tests compare the
// block-LDL^T factorisation with SuperLU directly.
int plfem_debug_solve(plfem_problem* pb, double sigma, const double* b, double* x, int refine) {
  if (!pb || !pb->ctx) return PLFEM_ERR_INVALID;
  plfem_ctx* ctx = pb->ctx;
  return guarded(ctx, [&] {
    need(b && x, "NULL argument");
    if (!pb->work || !pb->work->ready || pb->work->d_vals.p == nullptr) throw StatusError(PLFEM_ERR_NOT_READY, "run plfem_solve_modes first");
    SolveWork& W = *pb->work;
    (void)sigma;   // the factors and W.d_sigma belong to the last solve
    PLFEM_CUDA(cudaSetDevice(ctx->device));
    const int64_t n = W.dpat.n, m = 2 * n;
    std::vector<double> hb(m), hx(m);
    for (int64_t r = 0; r < n; ++r) { const int32_t ip = pb->plan.perm[r]; hb[2 * r] = b[ip]; hb[2 * r + 1] = b[n + ip]; }
    DevBuf<double> db, dx, dt, dd;
    db.upload(ctx, hb); dx.alloc(ctx, m); dt.alloc(ctx, m); dd.alloc(ctx, m);
    const int steps = refine >= 100 ? refine - 100 : refine;     // (100 + r is accepted for old callers)
    run_solve(ctx, W.dplan, db.p, dx.p);
    for (int it = 0; it < steps; ++it) {
      launch_resid_k(ctx, W.dpat, W.d_vals.p, W.d_sigma.p, dx.p, db.p, dt.p);
      run_solve(ctx, W.dplan, dt.p, dd.p);
      launch_axpy(ctx, dx.p, dd.p, m);
    }
    dx.download(hx.data(), m);
    PLFEM_CUDA(stream_wait(ctx->stream));
    for (int64_t r = 0; r < n; ++r) { const int32_t ip = pb->plan.perm[r]; x[ip] = hx[2 * r]; x[n + ip] = hx[2 * r + 1]; }
  });
}

int plfem_plan_sizes(plfem_problem* pb, int32_t leaf_nodes, int32_t max_sn_nodes, int64_t sizes[6]) {
  if (!pb || !sizes) return PLFEM_ERR_INVALID;
  return guarded(pb->ctx, [&] {
    ensure_plan(pb, leaf_nodes, max_sn_nodes, false);
    const FrontPlan& P = pb->plan;
    sizes[0] = P.n; sizes[1] = P.nfronts; sizes[2] = P.nlevels; sizes[3] = (int64_t)P.strct.size();
    sizes[4] = (int64_t)P.cmap.size(); sizes[5] = (int64_t)P.child.size();
  });
}

int plfem_plan_export(plfem_problem* pb, int32_t* perm, int32_t* first, int32_t* s, int32_t* parent, int32_t* level,
                      int32_t* sptr, int32_t* strct, int32_t* cmap_ptr, int32_t* cmap, int64_t* foff) {
  if (!pb) return PLFEM_ERR_INVALID;
  return guarded(pb->ctx, [&] {
    if (!pb->plan_ready) throw StatusError(PLFEM_ERR_NOT_READY, "call plfem_plan_sizes first");
    const FrontPlan& P = pb->plan;
    auto cp = [](const auto& v, auto* dst) { if (dst) std::copy(v.begin(), v.end(), dst); };
    cp(P.perm, perm); cp(P.first, first); cp(P.s, s); cp(P.parent, parent); cp(P.level, level); cp(P.sptr, sptr);
    cp(P.strct, strct); cp(P.cmap_ptr, cmap_ptr); cp(P.cmap, cmap); cp(P.foff, foff);
  });
}

int plfem_host_alloc(size_t bytes, void** out) {
  if (!out) return PLFEM_ERR_INVALID;
  *out = nullptr;
  return cudaHostAlloc(out, std::max<size_t>(bytes, 1), cudaHostAllocPortable) == cudaSuccess ? PLFEM_OK : PLFEM_ERR_CUDA;
}

void plfem_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

// test hook: dense symmetric eigensolver used at Lanczos restarts
int plfem_debug_symeig(int32_t n, double* a, double* w) {
  if (!a || !w || n < 1) return PLFEM_ERR_INVALID;
  std::vector<double> A(a, a + (size_t)n * n), W;
  symmetric_eigen(n, A, W);
  std::copy(A.begin(), A.end(), a);
  std::copy(W.begin(), W.end(), w);
  return PLFEM_OK;
}

// test hook: eigenvalues + the last p rows of the eigenvector matrix (what a convergence check reads); tail is p*n, column-major
int plfem_debug_symeig_tail(int32_t n, const double* a, double* w, int32_t p, double* tail) {
  if (!a || !w || !tail || n < 1 || p < 1 || p > n) return PLFEM_ERR_INVALID;
  std::vector<double> A(a, a + (size_t)n * n), W, T;
  symmetric_eigen_tail(n, A, W, p, T);
  std::copy(W.begin(), W.end(), w);
  std::copy(T.begin(), T.end(), tail);
  return PLFEM_OK;
}

}  // extern "C"
