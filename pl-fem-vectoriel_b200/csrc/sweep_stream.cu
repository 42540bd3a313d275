// Sweeps over the BOTTOM of the elimination forest: TMA-streamed subtrees, one warp per subtree.
//
// Replaces, for the fronts it covers, the triangular solves of SuperLU inside scipy eigsh (solver_fem.py:197).
//
// Where the time of a sweep went (round 1, ncu): the bottom levels of a forest are ten thousand small fronts
// (leaves 41 x 76 unknowns, the separators above them 13 x 66); a 256-thread CTA per front spent ~700 warp
// instructions on a 7 KB panel and a level of them took 50-60 us for 30 MB.  Here a bottom subtree — a maximal
// subtree whose pivots plus the update set of its root fit a local vector of ST_NLOC unknowns, typically 2-4 leaves
// with the separators between them, ~100 KB of factor — is ONE task for ONE warp (a 32-thread CTA):
//   * everything static about the subtree — the panels of its fronts in processing order, with the local index
//     maps between them — was laid out at pack time as one contiguous stream of 4 KB chunks; lane 0 keeps
//     ST_NS chunks in flight with cp.async.bulk (TMA, completion on an mbarrier per ring stage), starting before
//     the kernel's dependency wait, so no panel load is ever on the critical path and a front costs no address
//     arithmetic at all;
//   * the right-hand side of the subtree's (contiguous) pivot range and the update vector of its root live in
//     shared memory: a front reads its pivot part in place, multiplies it with the streamed panel
//     [F11^-1 ; W^T] and subtracts the update part from the local positions of its ancestors (forward), or gathers
//     its ancestors' solution from local positions (backward) — no update stacks, no child gather tables, no
//     global round trip between the fronts of a subtree;
//   * the root's update vector goes to the global update pool, where the per-level kernels of the fronts above
//     pick it up exactly like that of any other child.
// Which fronts are covered depends on the plan of their own design only, and every sum has a fixed order: a design
// gives bit-identical results alone and inside a forest.
#include "common.h"

#include <cstring>

namespace plfem {

namespace {

constexpr int CHD = ST_CHUNK_DOUBLES;   // doubles per chunk
constexpr int NS = 2;                   // ring stages per warp
constexpr int NLOC = ST_NLOC;           // local unknowns (pivots of the subtree + update set of its root)
constexpr int MAXF = 32;                // fronts per subtree: one descriptor per lane
constexpr int RB = 128;                 // rows of a forward row block (4 rows per lane)
constexpr int MAX_PIV_UNKNOWNS = 128;   // largest pivot block (unknowns)

__host__ __device__ inline int lm_item(int u) { return ((u + 7) / 8) * 2; }   // doubles holding u uint16 node indices, 16-byte multiple

// Position inside a subtree's stream.  An item (an index map, a panel column) never straddles a chunk boundary: if it
// does not fit into the rest of the current chunk it starts the next one.  The packer, the host layout and the
// consumer follow this one rule.
struct Cursor {
  int chunk, pos;
  __host__ __device__ int place(int sz) {
    if (pos + sz > CHD) { ++chunk; pos = 0; }
    const int o = chunk * CHD + pos;
    pos += sz;
    return o;
  }
  void place_n(int n, int sz) {          // n items of equal size, O(1)
    if (n <= 0) return;
    const int n0 = (CHD - pos) / sz;
    if (n <= n0) { pos += n * sz; return; }
    n -= n0;
    const int per = CHD / sz;
    chunk += 1 + (n - 1) / per;
    pos = ((n - 1) % per + 1) * sz;
  }
  int chunks() const { return chunk + (pos > 0 ? 1 : 0); }
};

// ---- PTX: mbarrier + bulk async copy (TMA) -------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(__cvta_generic_to_global(src)), "r"(bytes), "r"(smem_u32(b))
               : "memory");
}
// the rest of a task's sub-stream on its way into L2 while the ring holds only NSW chunks: the refills then cost an L2 round
// trip instead of a DRAM one (a 32 KB sub-stream through an 8 KB ring took 4-6 us, four DRAM latencies)
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(__cvta_generic_to_global(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
               : "=r"(ok)
               : "r"(smem_u32(b)), "r"(parity)
               : "memory");
  return ok != 0;
}
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// The warp's view of its stream: a ring of NST chunks, refilled by lane 0 as soon as the warp has left a chunk.
template <int NST>
struct PipeT {
  double* ring; uint64_t* bar; const double* src; int32_t* status;
  int nchunks, cur, pos, lane;
  int base = 0;                                       // chunks this ring has carried before (persistent CTAs: stage and parity run on)
  bool dead = false;                                  // a wait timed out: stop waiting, the status flag reports it
  __device__ __forceinline__ int stage() const { return (base + cur) % NST; }
  __device__ __forceinline__ void issue(int c) {
    const int s = (base + c) % NST;
    mbar_expect_tx(bar + s, CHD * 8);
    bulk_g2s(ring + s * CHD, src + (int64_t)c * CHD, CHD * 8, bar + s);
  }
  __device__ __forceinline__ void start(bool prefetch_rest = false) {
    if (lane == 0) {
      for (int c = 0; c < NST && c < nchunks; ++c) issue(c);
      if (prefetch_rest && nchunks > NST) bulk_prefetch_l2(src + (int64_t)NST * CHD, (uint32_t)(nchunks - NST) * CHD * 8);
    }
    cur = -1; pos = CHD;
  }
  __device__ __forceinline__ void advance() {
    if (cur >= 0) {
      __syncwarp();                                   // every lane is done with the chunk being left
      if (lane == 0 && cur + NST < nchunks) issue(cur + NST);
    }
    ++cur; pos = 0;
    uint64_t* b = bar + (base + cur) % NST;
    const uint32_t parity = ((base + cur) / NST) & 1;
    unsigned spins = 0;
    while (!dead && !mbar_try_wait(b, parity))
      if (++spins > (1u << 18)) { atomicExch(status + 1, 2); dead = true; }   // report instead of hanging the GPU
  }
  __device__ __forceinline__ const double* place(int sz) {
    if (pos + sz > CHD) advance();
    const double* p = ring + stage() * CHD + pos;
    pos += sz;
    return p;
  }
  __device__ __forceinline__ void drain() {           // no copy may be in flight when the CTA exits
    while (cur + 1 < nchunks) advance();
  }
};

using Pipe = PipeT<NS>;

// Optional stage clock of the dataflow forward sweep (PLFEM_TRACE_FILE, see profile_work in api.cu): eight %globaltimer
// stamps per task.  A null pointer (always, outside that measurement) costs one uniform load per task.
__device__ long long* g_sweep_trace = nullptr;
__device__ __forceinline__ void trace_stamp(long long* tr, int ti, int stage, int lane) {
  if (tr && lane == 0) {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    tr[(int64_t)ti * 8 + stage] = t;
  }
}

template <int NR>
struct StreamSmem {
  alignas(128) double ring[NS * CHD];
  alignas(16) double loc[NLOC * NR];     // [local unknown][rhs]
  alignas(16) uint16_t lm[NLOC / 2];     // local node index of the current front's update nodes
  alignas(8) uint64_t bar[NS];
};

template <int NR>
__device__ __forceinline__ void ld_loc(const double* loc, int idx, double (&v)[NR]) {
  if constexpr (NR % 2 == 0) {
    const double2* p = reinterpret_cast<const double2*>(loc + idx * NR);
#pragma unroll
    for (int r = 0; r < NR / 2; ++r) { const double2 t = p[r]; v[2 * r] = t.x; v[2 * r + 1] = t.y; }
  } else {
#pragma unroll
    for (int r = 0; r < NR; ++r) v[r] = loc[idx * NR + r];
  }
}
template <int NR>
__device__ __forceinline__ void ldg_v(const double* base, int64_t idx, double (&v)[NR]) {
  if constexpr (NR % 2 == 0) {
    const double2* p = reinterpret_cast<const double2*>(base + idx * NR);
#pragma unroll
    for (int r = 0; r < NR / 2; ++r) { const double2 t = p[r]; v[2 * r] = t.x; v[2 * r + 1] = t.y; }
  } else {
#pragma unroll
    for (int r = 0; r < NR; ++r) v[r] = base[idx * NR + r];
  }
}
// the same through L2 only (ld.global.cg): values another CTA of the SAME launch has just written (dataflow sweeps)
template <int NR>
__device__ __forceinline__ void ldcg_v(const double* base, int64_t idx, double (&v)[NR]) {
  if constexpr (NR % 2 == 0) {
    const double2* p = reinterpret_cast<const double2*>(base + idx * NR);
#pragma unroll
    for (int r = 0; r < NR / 2; ++r) { const double2 t = __ldcg(p + r); v[2 * r] = t.x; v[2 * r + 1] = t.y; }
  } else {
#pragma unroll
    for (int r = 0; r < NR; ++r) v[r] = __ldcg(base + idx * NR + r);
  }
}
// predicated vector load: zeros when idx < 0.  Written so that a run of them compiles into loads issued back to back — ONE
// memory round trip for the batch instead of one per load (measured with the stage clock: the chained version spent
// 2.5-4 us gathering and 1.6 us more before the first FMA of every task)
template <int NR>
__device__ __forceinline__ void ldcg_opt(const double* base, int64_t idx, double (&v)[NR]) {
  const bool ok = idx >= 0;
  if constexpr (NR % 2 == 0) {
    const double2* p = reinterpret_cast<const double2*>(base + (ok ? idx : 0) * NR);
#pragma unroll
    for (int r = 0; r < NR / 2; ++r) {
      double2 t = make_double2(0.0, 0.0);
      if (ok) t = __ldcg(p + r);
      v[2 * r] = t.x; v[2 * r + 1] = t.y;
    }
  } else {
#pragma unroll
    for (int r = 0; r < NR; ++r) v[r] = ok ? __ldcg(base + idx * NR + r) : 0.0;
  }
}

// n consecutive vectors src[g0 ..) into shared memory, four loads of a lane in flight at a time
template <int NR>
__device__ __forceinline__ void load_run(const double* src, int64_t g0, int n, double* dst, int lane) {
  for (int i0 = 0; i0 < n; i0 += 128) {
    double v[4][NR];
#pragma unroll
    for (int q = 0; q < 4; ++q) { const int i = i0 + lane + 32 * q; ldcg_opt<NR>(src, i < n ? g0 + i : -1, v[q]); }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int i = i0 + lane + 32 * q;
      if (i >= n) continue;
#pragma unroll
      for (int r = 0; r < NR; ++r) dst[i * NR + r] = v[q][r];
    }
  }
}
__device__ __forceinline__ int ld_acquire(const int32_t* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// lane 0 waits until counter c has reached `need`; gives up (and says so) instead of hanging the GPU
__device__ __forceinline__ void wait_counter(const int32_t* c, int need, int lane, int32_t* status) {
  if (lane == 0) {
    unsigned spins = 0;
    while (ld_acquire(c) < need) {
      __nanosleep(64);
      if (++spins > (1u << 22)) { atomicExch(status + 1, 3); break; }
    }
  }
  __syncwarp();
}
// every lane's stores first, then one increment
__device__ __forceinline__ void signal_counter(int32_t* c, int lane) {
  __threadfence();
  __syncwarp();
  if (lane == 0) atomicAdd(c, 1);
}

template <int NR>
__device__ __forceinline__ void stg_v(double* base, int64_t idx, const double (&v)[NR]) {
  if constexpr (NR % 2 == 0) {
    double2* p = reinterpret_cast<double2*>(base + idx * NR);
#pragma unroll
    for (int r = 0; r < NR / 2; ++r) p[r] = make_double2(v[2 * r], v[2 * r + 1]);
  } else {
#pragma unroll
    for (int r = 0; r < NR; ++r) base[idx * NR + r] = v[r];
  }
}

// The multiply of every sweep task: acc(row, rhs) += panel(row, c) * y(c, rhs) over n columns of a chunk in shared memory.
// A lane's FMAs on one accumulator are a dependent chain, and the FP64 pipe of this part returns a result only after tens of
// cycles (stage clock: 36 ns per column with one accumulator set, whatever the row count): the columns are dealt round-robin
// to U accumulator sets, added pairwise at the end (the order is fixed by the task's layout, so results stay reproducible).
template <int R> struct AccSets { static constexpr int U = R == 1 ? 4 : 2; };

template <int NR, int R, int U, class YF>
__device__ __forceinline__ void fma_columns(const double* __restrict__ base, int ld, int n, const bool (&valid)[R], YF&& yf,
                                            double (&acc)[U][R][NR]) {
  int c = 0;
#pragma unroll 2
  for (; c + U <= n; c += U) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      double y[NR];
      yf(c + u, y);
#pragma unroll
      for (int q = 0; q < R; ++q) {
        const double m = valid[q] ? base[(c + u) * ld + 32 * q] : 0.0;
#pragma unroll
        for (int r = 0; r < NR; ++r) acc[u][q][r] = fma(m, y[r], acc[u][q][r]);
      }
    }
  }
#pragma unroll
  for (int u = 0; u < U - 1; ++u) {
    if (c + u >= n) break;
    double y[NR];
    yf(c + u, y);
#pragma unroll
    for (int q = 0; q < R; ++q) {
      const double m = valid[q] ? base[(c + u) * ld + 32 * q] : 0.0;
#pragma unroll
      for (int r = 0; r < NR; ++r) acc[u][q][r] = fma(m, y[r], acc[u][q][r]);
    }
  }
}
template <int NR, int R, int U>
__device__ __forceinline__ void zero_sets(double (&acc)[U][R][NR]) {
#pragma unroll
  for (int u = 0; u < U; ++u)
#pragma unroll
    for (int q = 0; q < R; ++q)
#pragma unroll
      for (int r = 0; r < NR; ++r) acc[u][q][r] = 0.0;
}
template <int NR, int R, int U>
__device__ __forceinline__ void fold_sets(const double (&acc)[U][R][NR], int q, double (&out)[NR]) {
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    if constexpr (U == 4) out[r] = (acc[0][q][r] + acc[1][q][r]) + (acc[2][q][r] + acc[3][q][r]);
    else out[r] = acc[0][q][r] + acc[1][q][r];
  }
}

// forward, one row block (<= 128 rows, R per lane) of one front: acc = block * y1, y1 = loc[p0 ..) in place
template <int NR, int R>
__device__ __forceinline__ void fwd_block(Pipe& pp, StreamSmem<NR>& sm, int s2, int p0, int r0, int nrb, double* __restrict__ out,
                                          int64_t gpiv) {
  const int lane = pp.lane;
  const int ld = (nrb + 3) & ~3;
  bool valid[R];
#pragma unroll
  for (int q = 0; q < R; ++q) valid[q] = lane + 32 * q < nrb;
  constexpr int U = AccSets<R>::U;
  double accs[U][R][NR];
  zero_sets<NR, R, U>(accs);
  for (int k = 0; k < s2;) {
    int n = (CHD - pp.pos) / ld;
    if (n == 0) { pp.advance(); continue; }
    n = min(n, s2 - k);
    const double* base = pp.ring + pp.stage() * CHD + pp.pos + lane;
    const double* yv = sm.loc + (p0 + k) * NR;
    fma_columns<NR, R, U>(base, ld, n, valid, [&](int c, double (&y)[NR]) { ld_loc<NR>(yv, c, y); }, accs);
    pp.pos += n * ld; k += n;
  }
#pragma unroll
  for (int q = 0; q < R; ++q) {
    if (!valid[q]) continue;
    double acc[1][NR];
    fold_sets<NR, R, U>(accs, q, acc[0]);
    const int row = r0 + lane + 32 * q;
    if (row < s2) {
      stg_v<NR>(out, gpiv + row, acc[0]);                        // z1 = F11^-1 y1
    } else {
      const int j = row - s2;
      double* t = sm.loc + (2 * (int)sm.lm[j >> 1] + (j & 1)) * NR;   // the ancestor's local position: y -= W^T y1
#pragma unroll
      for (int r = 0; r < NR; ++r) t[r] -= acc[0][r];
    }
  }
}

template <int NR, bool PDL>
__global__ void __launch_bounds__(32) stream_forward_kernel(const StreamSub* __restrict__ subs, const int4* __restrict__ fronts,
                                                            const double* __restrict__ stream, const double* __restrict__ rhs,
                                                            double* __restrict__ out, double* __restrict__ upd, int32_t* status,
                                                            const uint8_t* __restrict__ active) {
  __shared__ StreamSmem<NR> sm;
  const int lane = threadIdx.x;
  if (PDL) griddep_launch_dependents();
  const StreamSub sb = subs[blockIdx.x];
  if (active && !active[sb.g0 >> 1]) return;        // a design of the forest that takes no part in this solve (refinement it does not need)
  if (lane == 0) {
    for (int s = 0; s < NS; ++s) mbar_init(sm.bar + s, 1);
    fence_mbar_init();
  }
  __syncwarp();
  Pipe pp{sm.ring, sm.bar, stream + sb.foff, status, sb.fchunks, -1, CHD, lane};
  pp.start();                                                     // static: the stream is on its way before the wait
  int4 fd = make_int4(0, 0, 0, 0);
  if (lane < sb.nfronts) fd = fronts[sb.front0 + lane];
  if (PDL) griddep_wait();
  // local vector: right-hand side of the subtree's pivots, zeros for the root's update set
  load_run<NR>(rhs, sb.g0, sb.nI, sm.loc, lane);
  for (int i = sb.nI + lane; i < sb.nI + sb.next; i += 32)
#pragma unroll
    for (int r = 0; r < NR; ++r) sm.loc[i * NR + r] = 0.0;
  __syncwarp();
  for (int i = 0; i < sb.nfronts; ++i) {
    const int su = __shfl_sync(0xffffffffu, fd.x, i), p0 = __shfl_sync(0xffffffffu, fd.y, i);
    const int s2 = su & 0xffff, u2 = su >> 16, rows = s2 + u2;
    if (u2 > 0) {
      const uint16_t* lmp = reinterpret_cast<const uint16_t*>(pp.place(lm_item(u2 >> 1)));
      for (int j = lane; j < (u2 >> 1); j += 32) sm.lm[j] = lmp[j];
      __syncwarp();
    }
    for (int r0 = 0; r0 < rows; r0 += RB) {
      const int nrb = min(RB, rows - r0);
      const int64_t gpiv = (int64_t)sb.g0 + p0;
      switch ((nrb + 31) >> 5) {
        case 1: fwd_block<NR, 1>(pp, sm, s2, p0, r0, nrb, out, gpiv); break;
        case 2: fwd_block<NR, 2>(pp, sm, s2, p0, r0, nrb, out, gpiv); break;
        case 3: fwd_block<NR, 3>(pp, sm, s2, p0, r0, nrb, out, gpiv); break;
        default: fwd_block<NR, 4>(pp, sm, s2, p0, r0, nrb, out, gpiv); break;
      }
      __syncwarp();                                               // the local vector is read by other lanes next
    }
  }
  // the root's update vector: what the subtree sends to the front above it
  for (int e = lane; e < sb.next; e += 32) {
    double v[NR];
    ld_loc<NR>(sm.loc, sb.nI + e, v);
    stg_v<NR>(upd, (int64_t)sb.uoff + e, v);
  }
  pp.drain();
}

// backward, one front: x1 = z1 - W x2 with x2 gathered from local positions; W row-major [u2][s2p]
template <int NR, int R>
__device__ __forceinline__ void bwd_front(Pipe& pp, StreamSmem<NR>& sm, int s2, int u2, int p0) {
  const int lane = pp.lane;
  const int s2p = (s2 + 3) & ~3;
  bool valid[R];
#pragma unroll
  for (int q = 0; q < R; ++q) valid[q] = lane + 32 * q < s2;
  constexpr int U = AccSets<R>::U;
  double accs[U][R][NR];
  zero_sets<NR, R, U>(accs);
  for (int j = 0; j < u2;) {
    int n = (CHD - pp.pos) / s2p;
    if (n == 0) { pp.advance(); continue; }
    n = min(n, u2 - j);
    const double* base = pp.ring + pp.stage() * CHD + pp.pos + lane;
    fma_columns<NR, R, U>(base, s2p, n, valid, [&](int c, double (&x)[NR]) {
      const int jj = j + c;
      ld_loc<NR>(sm.loc, 2 * (int)sm.lm[jj >> 1] + (jj & 1), x);
    }, accs);
    pp.pos += n * s2p; j += n;
  }
#pragma unroll
  for (int q = 0; q < R; ++q) {
    if (!valid[q]) continue;
    double a[NR];
    fold_sets<NR, R, U>(accs, q, a);
    double* t = sm.loc + (p0 + lane + 32 * q) * NR;
#pragma unroll
    for (int r = 0; r < NR; ++r) t[r] -= a[r];
  }
}

template <int NR, bool PDL>
__global__ void __launch_bounds__(32) stream_backward_kernel(const StreamSub* __restrict__ subs, const int4* __restrict__ fronts,
                                                             const double* __restrict__ stream, const int32_t* __restrict__ strct,
                                                             double* __restrict__ x, int32_t* status, const uint8_t* __restrict__ active) {
  __shared__ StreamSmem<NR> sm;
  const int lane = threadIdx.x;
  if (PDL) griddep_launch_dependents();
  const StreamSub sb = subs[blockIdx.x];
  if (active && !active[sb.g0 >> 1]) return;
  if (lane == 0) {
    for (int s = 0; s < NS; ++s) mbar_init(sm.bar + s, 1);
    fence_mbar_init();
  }
  __syncwarp();
  Pipe pp{sm.ring, sm.bar, stream + sb.boff, status, sb.bchunks, -1, CHD, lane};
  pp.start();
  int4 fd = make_int4(0, 0, 0, 0);
  if (lane < sb.nfronts) fd = fronts[sb.front0 + lane];
  // static too: where the root's update unknowns live in the global solution vector
  constexpr int NE = (NLOC + 31) / 32;
  int64_t xo[NE];
#pragma unroll
  for (int t = 0; t < NE; ++t) {
    const int e = lane + 32 * t;
    xo[t] = e < sb.next ? 2 * (int64_t)strct[sb.soff + (e >> 1)] + (e & 1) : -1;
  }
  if (PDL) griddep_wait();
  load_run<NR>(x, sb.g0, sb.nI, sm.loc, lane);                    // z of the subtree's pivots (forward sweep)
  constexpr int NEH = (NE + 1) / 2;                               // x of the ancestors above the subtree (final): two batches
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    double v[NEH][NR];
#pragma unroll
    for (int t = 0; t < NEH; ++t) ldcg_opt<NR>(x, h * NEH + t < NE ? xo[(h * NEH + t) % NE] : -1, v[t]);
#pragma unroll
    for (int t = 0; t < NEH; ++t) {
      const int tt = h * NEH + t;
      if (tt >= NE || xo[tt % NE] < 0) continue;
      const int e = lane + 32 * tt;
#pragma unroll
      for (int r = 0; r < NR; ++r) sm.loc[(sb.nI + e) * NR + r] = v[t][r];
    }
  }
  __syncwarp();
  for (int i = sb.nfronts - 1; i >= 0; --i) {                     // the root first
    const int su = __shfl_sync(0xffffffffu, fd.x, i), p0 = __shfl_sync(0xffffffffu, fd.y, i);
    const int s2 = su & 0xffff, u2 = su >> 16;
    if (u2 == 0) continue;
    const uint16_t* lmp = reinterpret_cast<const uint16_t*>(pp.place(lm_item(u2 >> 1)));
    for (int j = lane; j < (u2 >> 1); j += 32) sm.lm[j] = lmp[j];
    __syncwarp();
    switch ((s2 + 31) >> 5) {
      case 1: bwd_front<NR, 1>(pp, sm, s2, u2, p0); break;
      case 2: bwd_front<NR, 2>(pp, sm, s2, u2, p0); break;
      case 3: bwd_front<NR, 3>(pp, sm, s2, u2, p0); break;
      default: bwd_front<NR, 4>(pp, sm, s2, u2, p0); break;
    }
    __syncwarp();
  }
  for (int i = lane; i < sb.nI; i += 32) {
    double v[NR];
    ld_loc<NR>(sm.loc, i, v);
    stg_v<NR>(x, (int64_t)sb.g0 + i, v);
  }
  pp.drain();
}

// ---- pack: the fronts of the bottom subtrees go from the front pool into the two streams --------------------------
__global__ void __launch_bounds__(256) stream_pack_kernel(const StreamPackRec* __restrict__ recs, const StreamSub* __restrict__ subs,
                                                          const int32_t* __restrict__ s_of, const int32_t* __restrict__ sptr,
                                                          const int64_t* __restrict__ foff, const double* __restrict__ pool,
                                                          const uint16_t* __restrict__ lmaps, double* __restrict__ sfwd,
                                                          double* __restrict__ sbwd) {
  __shared__ int offF[(NLOC + RB - 1) / RB * 128];
  __shared__ int offB[NLOC];
  __shared__ int lmF, lmB;
  const StreamPackRec rec = recs[blockIdx.x];
  const int f = rec.f, tid = threadIdx.x;
  const int s2 = 2 * s_of[f], u = sptr[f + 1] - sptr[f], u2 = 2 * u, rows = s2 + u2;
  const int s2p = (s2 + 3) & ~3;
  const int64_t ld = rows;
  if (tid == 0) {
    Cursor c{rec.fchunk, rec.fpos};
    lmF = u ? c.place(lm_item(u)) : 0;
    for (int r0 = 0, bi = 0; r0 < rows; r0 += RB, ++bi) {
      const int ldb = (min(RB, rows - r0) + 3) & ~3;
      for (int k = 0; k < s2; ++k) offF[bi * 128 + k] = c.place(ldb);
    }
    Cursor d{rec.bchunk, rec.bpos};
    lmB = u ? d.place(lm_item(u)) : 0;
    for (int j = 0; j < u2; ++j) offB[j] = d.place(s2p);
  }
  __syncthreads();
  const double* src = pool + foff[f];
  double* F = sfwd + subs[rec.sub].foff;
  double* Bk = sbwd + subs[rec.sub].boff;
  if (u) {
    uint16_t* a = reinterpret_cast<uint16_t*>(F + lmF);
    uint16_t* b = reinterpret_cast<uint16_t*>(Bk + lmB);
    for (int i = tid; i < lm_item(u) * 4; i += 256) {
      const uint16_t v = i < u ? lmaps[rec.lm_off + i] : (uint16_t)0;
      a[i] = v; b[i] = v;
    }
  }
  for (int r0 = 0, bi = 0; r0 < rows; r0 += RB, ++bi) {
    const int nrb = min(RB, rows - r0), ldb = (nrb + 3) & ~3;
    for (int idx = tid; idx < s2 * ldb; idx += 256) {
      const int k = idx / ldb, i = idx - k * ldb;
      F[offF[bi * 128 + k] + i] = i < nrb ? src[k * ld + r0 + i] : 0.0;
    }
  }
  // W(j, c) = src[c * ld + s2 + j]: read with j fastest (coalesced), write row-major [j][s2p]
  for (int idx = tid; idx < u2 * s2; idx += 256) {
    const int c = idx / u2, j = idx - c * u2;
    Bk[offB[j] + c] = src[c * ld + s2 + j];
  }
  for (int idx = tid; idx < u2 * (s2p - s2); idx += 256) {
    const int j = idx / (s2p - s2), c = s2 + idx % (s2p - s2);
    Bk[offB[j] + c] = 0.0;
  }
}

// ================================================================================================================
// Fronts ABOVE the bottom subtrees: tasks of FOUR warps (a 128-thread CTA), one thread per row (forward) or per pivot
// column (backward).
//
// Why four warps: one warp of this part issues an FP64 FMA (or a shared load) only every 4-5 cycles, whatever the number of
// independent accumulators (scripts/micro/panel_fma.cu: 26 cycles per panel column with 32 rows, 65-75 with 128 rows), so
// a one-warp task of 32 KB spent 4-5 us in its multiply — more than in all its memory round trips together (stage clock,
// scripts/sweep_trace.py) — and held 170 registers per thread for its 4 x 4 accumulators, which capped an SM at 11 warps.
// Here every thread owns ONE row: a forward task is a slab of <= 128 rows of a front's left block column
// [F11^-1 ; W^T], a backward task a slab of <= 128 ... 256 update unknowns against all pivot columns of W; warp w streams
// its own quarter (rows / pivot columns 32 w ...) from its own chunk-aligned sub-stream through its own ring of bulk copies
// (started before the dependency wait), the four warps share only the assembled contraction vector, and the task takes a
// quarter of the time with half the registers (24 warps per SM).
// Backward tasks split the CONTRACTION: no two tasks gather the same ancestors' values (the first version split the pivot
// columns: sixteen tasks of a front with 3 800 update unknowns each gathered all of them, and each ran for 50 us).  The slabs
// of a front leave partial sums in `part`; whichever arrives last adds them IN SLAB ORDER — the result does not depend on
// who that is — and subtracts from z1.
constexpr int KSMAX = 256;  // contraction entries of a backward task (a slab of the update set)
constexpr int TW = 4;       // warps per task
constexpr int NSW = 2;      // ring stages per warp: 8 KB in flight per warp, 32 KB per task

struct RhsView {            // interleaved vectors of a forward sweep: right-hand sides in, pivot solutions out, update pool
  const double* rhs; double* out; double* upd;
};

using PipeW = PipeT<NSW>;

template <int NR>
struct LevelFwdSmem {
  alignas(128) double ring[TW][NSW * CHD];
  alignas(16) double cv[MAX_PIV_UNKNOWNS * NR];   // assembled pivot part of the right-hand sides
  alignas(8) uint64_t bar[TW][NSW];
  int ticket;
};
template <int NR>
struct LevelBwdSmem {
  alignas(128) double ring[TW][NSW * CHD];
  alignas(16) double cv[KSMAX * NR];              // gathered x2 of the task's slab of the update set
  alignas(8) uint64_t bar[TW][NSW];
  int ticket, last;
};

// uniform items of size sz, starting chunk-aligned: item i of a sub-stream
__host__ __device__ inline int64_t item_off(int i, int sz) {
  const int per = CHD / sz;
  return (int64_t)(i / per) * CHD + (i % per) * sz;
}
__host__ __device__ inline int item_chunks(int n, int sz) { return n == 0 ? 0 : (n + CHD / sz - 1) / (CHD / sz); }

// one warp's multiply: its 32 rows (or pivot columns) against n items of 32 doubles, y(c) from shared memory
template <int NR, class YF>
__device__ __forceinline__ void warp_multiply(PipeW& pp, int n, bool valid, YF&& yf, double (&out)[NR]) {
  const bool v1[1] = {valid};
  double accs[4][1][NR];
  zero_sets<NR, 1, 4>(accs);
  for (int k = 0; k < n;) {
    int m = (CHD - pp.pos) / 32;
    if (m == 0) { pp.advance(); continue; }
    m = min(m, n - k);
    const double* base = pp.ring + pp.stage() * CHD + pp.pos + pp.lane;
    fma_columns<NR, 1, 4>(base, 32, m, v1, [&](int c, double (&y)[NR]) { yf(k + c, y); }, accs);
    pp.pos += m * 32; k += m;
  }
  fold_sets<NR, 1, 4>(accs, 0, out);
}

template <int NR, bool PDL, bool FUSED>
__global__ void __launch_bounds__(32 * TW) level_forward_kernel(const LevelTask* __restrict__ tasks, const int32_t* __restrict__ gsrc,
                                                                const double* __restrict__ stream, const int32_t* __restrict__ cptr,
                                                                const int32_t* __restrict__ child, const int32_t* __restrict__ cmap_ptr,
                                                                const int32_t* __restrict__ cmap, const int32_t* __restrict__ sptr,
                                                                const int32_t* __restrict__ uoff, RhsView rv, int32_t* status, int32_t* sync,
                                                                const uint8_t* __restrict__ active, int ntasks) {
  __shared__ LevelFwdSmem<NR> sm;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (PDL) griddep_launch_dependents();
  if (lane == 0) {
    for (int s = 0; s < NSW; ++s) mbar_init(sm.bar[warp] + s, 1);
    fence_mbar_init();
  }
  __syncwarp();
  int ring_base = 0;                           // chunks this warp's ring has carried in earlier tasks of this CTA
  bool pdl_pending = PDL;
  long long* const tr = FUSED ? g_sweep_trace : nullptr;
  // Dataflow launch: a bounded number of PERSISTENT CTAs; each takes the next task by a ticket, in level order, so whatever
  // a running task waits for belongs to a task taken before it — by a CTA that is running — and progress does not depend
  // on how many CTAs the hardware has made resident or in which order.  (One CTA per task let thousands of waiting tasks of
  // one forest occupy the shared memory that runnable tasks of the other forests in flight needed.)
  for (bool once = true;; once = false) {
  int ti = blockIdx.x;
  if (FUSED) {
    __syncthreads();                           // the previous task of this CTA is done with the shared contraction vector
    if (tid == 0) sm.ticket = atomicAdd(sync, 1);
    __syncthreads();
    ti = sm.ticket;
    if (ti >= ntasks) break;
  } else if (!once) {
    break;
  }
  const LevelTask t = tasks[ti];
  // a design that takes no part in this solve: all fronts of its tree are skipped, so nobody waits for this task's signal
  if (active && !active[t.g0 >> 1]) continue;
  trace_stamp(tr, ti, 1, tid);
  const int nw = (t.n + 31) >> 5;              // warps with rows
  PipeW pp{sm.ring[warp], sm.bar[warp], stream + t.soff + (int64_t)warp * t.chunks * CHD, status, warp < nw ? t.chunks : 0, -1, CHD, lane, ring_base};
  if (warp < nw) pp.start(true);
  // static: where this thread's pivot row and its row of the slab receive their children's updates
  const int nf2 = t.s2 + t.u2;
  const int32_t* g1 = gsrc + t.goff;
  const int32_t* g2 = g1 + nf2;
  int i1 = -1, i2 = -1, j1 = -1, j2 = -1;
  double v[NR];
#pragma unroll
  for (int r = 0; r < NR; ++r) v[r] = 0.0;
  if (tid < t.s2) {          // the right-hand sides are input of the whole solve: final before the first sweep kernel started
    i1 = g1[tid]; i2 = g2[tid];
    ldg_v<NR>(rv.rhs, t.g0 + tid, v);
  }
  const int row = t.r0 + tid;
  const bool has_row = tid < t.n;
  if (has_row && row >= t.s2) { j1 = g1[row]; j2 = g2[row]; }
  if (pdl_pending) { griddep_wait(); pdl_pending = false; }
  trace_stamp(tr, ti, 2, tid);
  if (FUSED && t.need > 0) {                   // the children's update vectors are complete
    if (warp == 0) wait_counter(sync + 2 + t.dep, t.need, lane, status);
    __syncthreads();
  }
  trace_stamp(tr, ti, 3, tid);
  // assembled pivot part of the right-hand sides, fixed order (rhs + first child) + second child; both gathers in flight at once
  {
    double w1[NR], w2[NR];
    ldcg_opt<NR>(rv.upd, i1, w1); ldcg_opt<NR>(rv.upd, i2, w2);
    if (tid < t.s2) {
#pragma unroll
      for (int r = 0; r < NR; ++r) sm.cv[tid * NR + r] = (v[r] + w1[r]) + w2[r];
    }
  }
  __syncthreads();
  if (t.nch > 2) {          // rare (a separator that does not disconnect): the further children, one after the other
    for (int c = cptr[t.f] + 2; c < cptr[t.f + 1]; ++c) {
      const int ch = child[c];
      const int uc2 = 2 * (sptr[ch + 1] - sptr[ch]);
      const int32_t* cm = cmap + cmap_ptr[ch];
      for (int k = tid; k < uc2; k += 32 * TW) {
        const int prow = 2 * cm[k >> 1] + (k & 1);
        if (prow < t.s2) {
          double w[NR];
          ldcg_v<NR>(rv.upd, (int64_t)uoff[ch] + k, w);
#pragma unroll
          for (int r = 0; r < NR; ++r) sm.cv[prow * NR + r] += w[r];
        }
      }
      __syncthreads();
    }
  }
  trace_stamp(tr, ti, 4, tid);
  if (warp < nw) {
    // what the children send to this thread's update row: requested now, used after the panel has been multiplied
    double y1[NR], y2[NR], a[NR];
    ldcg_opt<NR>(rv.upd, j1, y1); ldcg_opt<NR>(rv.upd, j2, y2);
    warp_multiply<NR>(pp, t.s2, has_row, [&](int c, double (&y)[NR]) { ld_loc<NR>(sm.cv, c, y); }, a);
    if (warp == 0) trace_stamp(tr, ti, 7, tid);
    if (has_row) {
      if (row < t.s2) {
        stg_v<NR>(rv.out, t.g0 + row, a);                            // z1 = F11^-1 y1
      } else {
        double o[NR];
#pragma unroll
        for (int r = 0; r < NR; ++r) o[r] = (y1[r] + y2[r]) - a[r];   // first child, second child (fixed order), minus W^T y1
        stg_v<NR>(rv.upd, (int64_t)t.uoff + (row - t.s2), o);
      }
    }
  }
  if (t.nch > 2) {          // update rows of this task: contributions of the further children (after the first two, fixed order)
    __syncthreads();
    for (int c = cptr[t.f] + 2; c < cptr[t.f + 1]; ++c) {
      const int ch = child[c];
      const int uc2 = 2 * (sptr[ch + 1] - sptr[ch]);
      const int32_t* cm = cmap + cmap_ptr[ch];
      for (int k = tid; k < uc2; k += 32 * TW) {
        const int prow = 2 * cm[k >> 1] + (k & 1);
        if (prow >= t.s2 && prow >= t.r0 && prow < t.r0 + t.n) {
          double w[NR], o[NR];
          ldcg_v<NR>(rv.upd, (int64_t)uoff[ch] + k, w);
          ldcg_v<NR>(rv.upd, (int64_t)t.uoff + (prow - t.s2), o);
#pragma unroll
          for (int r = 0; r < NR; ++r) o[r] += w[r];
          stg_v<NR>(rv.upd, (int64_t)t.uoff + (prow - t.s2), o);
        }
      }
      __syncthreads();
    }
  }
  trace_stamp(tr, ti, 5, tid);
  if (FUSED && t.sig >= 0) {                   // every thread's stores first, then one increment
    __threadfence();
    __syncthreads();
    if (tid == 0) atomicAdd(sync + 2 + t.sig, 1);
  }
  trace_stamp(tr, ti, 6, tid);
  if (warp < nw) { pp.drain(); ring_base += t.chunks; }
  }
}

template <int NR, bool PDL, bool FUSED>
__global__ void __launch_bounds__(32 * TW) level_backward_kernel(const LevelTask* __restrict__ tasks, const double* __restrict__ stream,
                                                                 const int32_t* __restrict__ strct, double* __restrict__ x,
                                                                 double* __restrict__ part, int32_t* status, int32_t* sync, int nfronts,
                                                                 const uint8_t* __restrict__ active, int ntasks) {
  __shared__ LevelBwdSmem<NR> sm;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (PDL) griddep_launch_dependents();
  if (lane == 0) {
    for (int s = 0; s < NSW; ++s) mbar_init(sm.bar[warp] + s, 1);
    fence_mbar_init();
  }
  __syncwarp();
  int ring_base = 0;
  bool pdl_pending = PDL;
  long long* const tr = FUSED ? g_sweep_trace : nullptr;
  for (bool once = true;; once = false) {      // persistent CTAs, tickets in level order, top level first (see level_forward_kernel)
  int ti = blockIdx.x;
  if (FUSED) {
    __syncthreads();
    if (tid == 0) sm.ticket = atomicAdd(sync + 1, 1);
    __syncthreads();
    ti = sm.ticket;
    if (ti >= ntasks) break;
  } else if (!once) {
    break;
  }
  const LevelTask t = tasks[ti];
  if (active && !active[t.g0 >> 1]) continue;  // see level_forward_kernel
  trace_stamp(tr, ti, 1, tid);
  const int k0 = t.r0, kn = t.n, nk = t.nch;   // this task's slab of the update unknowns; slabs of the front
  const int nw = (t.s2 + 31) >> 5;             // warps with pivot columns
  PipeW pp{sm.ring[warp], sm.bar[warp], stream + t.soff + (int64_t)warp * t.chunks * CHD, status, warp < nw ? t.chunks : 0, -1, CHD, lane, ring_base};
  if (warp < nw) pp.start(true);
  const int32_t* st = strct + t.goff;          // the front's update set
  // static: positions of the slab's update unknowns in the solution vector (two per thread)
  int64_t xo[KSMAX / (32 * TW)];
#pragma unroll
  for (int q = 0; q < KSMAX / (32 * TW); ++q) {
    const int j = tid + 32 * TW * q;
    xo[q] = j < kn ? 2 * (int64_t)st[(k0 + j) >> 1] + ((k0 + j) & 1) : -1;
  }
  if (pdl_pending) { griddep_wait(); pdl_pending = false; }
  const bool has_col = tid < t.s2;
  // z1 of this thread's pivot (forward sweep): requested before the dependency wait, used at the very end
  double z[NR];
  ldcg_opt<NR>(x, has_col ? t.g0 + tid : -1, z);
  trace_stamp(tr, ti, 2, tid);
  if (FUSED && t.need > 0) {                   // the parent's unknowns are final
    if (warp == 0) wait_counter(sync + 2 + nfronts + t.dep, t.need, lane, status);
    __syncthreads();
  }
  trace_stamp(tr, ti, 3, tid);
  {
    double g[KSMAX / (32 * TW)][NR];           // the slab of x2: one batch of gathers
#pragma unroll
    for (int q = 0; q < KSMAX / (32 * TW); ++q) ldcg_opt<NR>(x, xo[q], g[q]);
#pragma unroll
    for (int q = 0; q < KSMAX / (32 * TW); ++q) {
      if (xo[q] < 0) continue;
#pragma unroll
      for (int r = 0; r < NR; ++r) sm.cv[(tid + 32 * TW * q) * NR + r] = g[q][r];
    }
  }
  __syncthreads();
  trace_stamp(tr, ti, 4, tid);
  double a[NR];
#pragma unroll
  for (int r = 0; r < NR; ++r) a[r] = 0.0;
  if (warp < nw) warp_multiply<NR>(pp, kn, has_col, [&](int c, double (&y)[NR]) { ld_loc<NR>(sm.cv, c, y); }, a);
  trace_stamp(tr, ti, 7, tid);
  bool finish = true;
  if (nk > 1) {
    const int s2p = (t.s2 + 3) & ~3;
    if (has_col) stg_v<NR>(part + ((int64_t)t.uoff + (int64_t)(k0 / t.pad[0]) * s2p) * NR, tid, a);    // pad[0]: slab length of this front
    __threadfence();
    __syncthreads();
    if (tid == 0) sm.last = (atomicAdd(sync + 2 + 2 * nfronts + t.f, 1) == nk - 1);
    __syncthreads();
    finish = sm.last != 0;
    if (finish) {
      __threadfence();                          // the other slabs' partial sums are visible now
#pragma unroll
      for (int r = 0; r < NR; ++r) a[r] = 0.0;
      const double* all = part + (int64_t)t.uoff * NR;
      if (has_col)
        for (int kk = 0; kk < nk; ++kk) {       // slab order, whoever finishes
          double w[NR];
          ldcg_v<NR>(all, (int64_t)kk * s2p + tid, w);
#pragma unroll
          for (int r = 0; r < NR; ++r) a[r] += w[r];
        }
    }
  }
  if (finish) {
    if (has_col) {
#pragma unroll
      for (int r = 0; r < NR; ++r) z[r] -= a[r];
      stg_v<NR>(x, t.g0 + tid, z);
    }
    trace_stamp(tr, ti, 5, tid);
    if (FUSED && t.sig >= 0) {
      __threadfence();
      __syncthreads();
      if (tid == 0) atomicAdd(sync + 2 + nfronts + t.sig, 1);
    }
  }
  trace_stamp(tr, ti, 6, tid);
  if (warp < nw) { pp.drain(); ring_base += t.chunks; }
  }
}

// pack of the level tasks: one CTA per task copies its slab from the front pool into the sub-streams of its warps
__global__ void __launch_bounds__(256) level_pack_kernel(const LevelTask* __restrict__ ftasks, int nf_tasks, const LevelTask* __restrict__ btasks,
                                                         const int64_t* __restrict__ foff, const double* __restrict__ pool,
                                                         double* __restrict__ lfwd, double* __restrict__ lbwd) {
  const bool fwd = (int)blockIdx.x < nf_tasks;
  const LevelTask t = fwd ? ftasks[blockIdx.x] : btasks[blockIdx.x - nf_tasks];
  const int f = t.f, tid = threadIdx.x;
  const int s2 = t.s2, u2 = t.u2;
  const int64_t ld = s2 + u2;
  const double* src = pool + foff[f];
  const int64_t wstride = (int64_t)t.chunks * CHD;      // one warp's sub-stream
  if (fwd) {
    // rows r0 .. r0 + n of the left block column: warp w owns rows 32 w ..., item k = column k of its 32 rows
    double* dst = lfwd + t.soff;
    const int nwr = ((t.n + 31) >> 5) * 32;
    for (int idx = tid; idx < s2 * nwr; idx += 256) {
      const int k = idx / nwr, i = idx - k * nwr;
      dst[(i >> 5) * wstride + item_off(k, 32) + (i & 31)] = i < t.n ? src[k * ld + t.r0 + i] : 0.0;
    }
  } else {
    // W(j, c) = src[c * ld + s2 + j] for the slab j = k0 .. k0 + kn: warp w owns pivot columns 32 w ..., item j - k0 = row j of its 32
    // columns; read with j fastest (coalesced)
    double* dst = lbwd + t.soff;
    const int k0 = t.r0, kn = t.n, ncw = ((s2 + 31) >> 5) * 32;
    for (int64_t idx = tid; idx < (int64_t)kn * ncw; idx += 256) {
      const int c = (int)(idx / kn), j = (int)(idx - (int64_t)c * kn);
      dst[(c >> 5) * wstride + item_off(j, 32) + (c & 31)] = c < s2 ? src[(int64_t)c * ld + s2 + k0 + j] : 0.0;
    }
  }
}

int stream_nloc() {
  static const int v = [] {
    const char* e = std::getenv("PLFEM_STREAM_NLOC");
    return e ? std::max(0, std::min(NLOC, atoi(e))) : NLOC;
  }();
  return v;
}

}  // namespace

// Choose the bottom subtrees, lay out their streams and index maps.  in_sub[f] = 1 for the fronts they cover.
void build_stream_plan(plfem_ctx* ctx, const FrontPlan& P, const std::vector<int32_t>& uoff, std::vector<uint8_t>& in_sub, StreamPlan& S) {
  const int nf = P.nfronts;
  in_sub.assign(nf, 0);
  S.n_subs = 0; S.n_fronts = 0;
  const int nloc = stream_nloc();
  if (nloc == 0 || nf == 0) return;
  std::vector<int32_t> size(nf, 1), nI(nf);
  for (int f = 0; f < nf; ++f) nI[f] = 2 * P.s[f];
  for (int f = 0; f < nf; ++f) {
    const int p = P.parent[f];
    if (p >= 0) { size[p] += size[f]; nI[p] += nI[f]; }            // post-order: children come before their parent
  }
  auto ok = [&](int f) { return nI[f] + 2 * (P.sptr[f + 1] - P.sptr[f]) <= nloc && size[f] <= MAXF; };
  std::vector<StreamSub> subs;
  std::vector<int4> fronts;
  std::vector<StreamPackRec> recs;
  std::vector<uint16_t> lmaps;
  int64_t fo = 0, bo = 0;     // chunks
  for (int r = 0; r < nf; ++r) {
    if (!ok(r)) continue;
    const int p = P.parent[r];
    if (p >= 0 && ok(p)) continue;                                  // nI + u2 and size grow towards the root: maximal subtrees
    const int f0 = r - size[r] + 1, node0 = P.first[f0];
    if (P.first[r] + P.s[r] - node0 != nI[r] / 2) throw StatusError(PLFEM_ERR_INTERNAL, "front plan: the pivots of a subtree are not contiguous");
    StreamSub sb{};
    sb.nfronts = size[r]; sb.front0 = (int32_t)fronts.size(); sb.g0 = 2 * node0; sb.nI = nI[r];
    sb.next = 2 * (P.sptr[r + 1] - P.sptr[r]); sb.uoff = uoff[r]; sb.soff = P.sptr[r];
    const int32_t* rs = P.strct.data() + P.sptr[r];
    const int ru = P.sptr[r + 1] - P.sptr[r];
    const size_t rec0 = recs.size();
    Cursor c{0, 0};
    for (int f = f0; f <= r; ++f) {
      in_sub[f] = 1;
      const int s2 = 2 * P.s[f], u = P.sptr[f + 1] - P.sptr[f], rows = s2 + 2 * u;
      StreamPackRec rec{};
      rec.f = f; rec.sub = (int32_t)subs.size(); rec.fchunk = c.chunk; rec.fpos = c.pos; rec.lm_off = (int32_t)lmaps.size();
      if (u) c.place(lm_item(u));
      for (int r0 = 0; r0 < rows; r0 += RB) c.place_n(s2, (std::min(RB, rows - r0) + 3) & ~3);
      recs.push_back(rec);
      fronts.push_back(make_int4(s2 | (2 * u) << 16, 2 * (P.first[f] - node0), 0, 0));
      // local node index of every update node: inside the subtree by its offset, above it by its place in the root's update set
      const int32_t* st = P.strct.data() + P.sptr[f];
      int q = 0;
      for (int j = 0; j < u; ++j) {
        const int t = st[j];
        if (t < node0 + nI[r] / 2) { lmaps.push_back((uint16_t)(t - node0)); continue; }
        while (q < ru && rs[q] < t) ++q;
        if (q >= ru || rs[q] != t) throw StatusError(PLFEM_ERR_INTERNAL, "front plan: an update node of a subtree is missing from its root's update set");
        lmaps.push_back((uint16_t)(nI[r] / 2 + q));
      }
    }
    sb.fchunks = c.chunks();
    Cursor d{0, 0};
    for (int f = r; f >= f0; --f) {
      const int s2 = 2 * P.s[f], u = P.sptr[f + 1] - P.sptr[f];
      StreamPackRec& rec = recs[rec0 + (f - f0)];
      rec.bchunk = d.chunk; rec.bpos = d.pos;
      if (u) { d.place(lm_item(u)); d.place_n(2 * u, (s2 + 3) & ~3); }
    }
    sb.bchunks = d.chunks();
    sb.foff = fo * CHD; sb.boff = bo * CHD;
    fo += sb.fchunks; bo += sb.bchunks;
    subs.push_back(sb);
  }
  S.n_subs = (int)subs.size(); S.n_fronts = (int)recs.size();
  if (S.n_subs == 0) return;
  S.subs.upload(ctx, subs); S.fronts.upload(ctx, fronts); S.recs.upload(ctx, recs); S.lmaps.upload(ctx, lmaps);
  S.sfwd.alloc(ctx, (size_t)std::max<int64_t>(fo, 1) * CHD); S.sbwd.alloc(ctx, (size_t)std::max<int64_t>(bo, 1) * CHD);
  S.fwd_doubles = fo * CHD; S.bwd_doubles = bo * CHD;
  PLFEM_CUDA(stream_wait(ctx->stream));       // the host vectors above are pageable and local
}

// Task lists of the fronts above the bottom subtrees, level by level, and the layout of their streams.  Forward tasks are
// stored bottom level first, backward tasks top level first: the order in which the dataflow launches hand them out.
void build_level_plan(plfem_ctx* ctx, const FrontPlan& P, const std::vector<int32_t>& uoff, const std::vector<uint8_t>& in_sub,
                      const std::vector<int32_t>& goff, StreamPlan& S) {
  std::vector<LevelTask> ft, bt;
  S.fptr.assign(P.nlevels + 1, 0); S.bptr.assign(P.nlevels + 1, 0);
  S.nfronts = P.nfronts;
  // slab sizes depend on the front alone (a design's arithmetic is the same alone and inside a forest)
  auto fwd_rows = [](int) { return 32 * TW; };        // a forward task: 128 rows, one per thread
  // backward: slabs of the contraction index of about 64 KB of W (all pivot columns x ks update unknowns; 32 / 64 / 128 KB measured
  // within 3 % of each other on cfg1 and cfg5)
  static const int slab_doubles = [] { const char* e = std::getenv("PLFEM_BWD_SLAB_KB"); return 128 * std::max(8, std::min(256, e ? atoi(e) : 64)); }();
  auto bwd_ks = [](int s2) { return std::max(32, std::min(KSMAX, (slab_doubles / ((s2 + 3) & ~3)) & ~31)); };
  std::vector<int32_t> nft(P.nfronts, 0), nbt(P.nfronts, 0);      // tasks per front
  for (int f = 0; f < P.nfronts; ++f) {
    if (in_sub[f]) continue;
    const int s2 = 2 * P.s[f], u2 = 2 * (P.sptr[f + 1] - P.sptr[f]);
    nft[f] = (s2 + u2 + fwd_rows(s2) - 1) / fwd_rows(s2);
    if (u2 > 0) nbt[f] = (u2 + bwd_ks(s2) - 1) / bwd_ks(s2);
  }
  std::vector<int32_t> fneed(P.nfronts, 0);                        // forward: tasks of the children above the subtrees
  for (int f = 0; f < P.nfronts; ++f)
    if (!in_sub[f] && P.parent[f] >= 0) fneed[P.parent[f]] += nft[f];
  int64_t fo = 0, bo = 0;     // chunks
  int64_t po = 0;             // partial sums of the backward slabs (vector entries)
  for (int l = 0; l < P.nlevels; ++l) {
    for (int q = P.lptr[l]; q < P.lptr[l + 1]; ++q) {
      const int f = P.lfront[q];
      if (in_sub[f]) continue;
      const int s2 = 2 * P.s[f], u2 = 2 * (P.sptr[f + 1] - P.sptr[f]), rows = s2 + u2;
      LevelTask t{};
      t.g0 = 2 * (int64_t)P.first[f]; t.s2 = s2; t.u2 = u2; t.f = f;
      // forward: slabs of 128 rows, four sub-streams (one per warp: s2 columns of 32 rows each)
      const int nr = fwd_rows(s2);
      t.goff = goff[f]; t.uoff = uoff[f]; t.nch = P.cptr[f + 1] - P.cptr[f];
      t.dep = f; t.need = fneed[f]; t.sig = P.parent[f];
      for (int r0 = 0; r0 < rows; r0 += nr) {
        t.r0 = r0; t.n = std::min(nr, rows - r0);
        t.chunks = item_chunks(s2, 32);                        // per warp
        t.soff = fo * CHD; fo += (int64_t)t.chunks * ((t.n + 31) >> 5);
        ft.push_back(t);
      }
    }
    S.fptr[l + 1] = (int32_t)ft.size();
  }
  for (int l = P.nlevels - 1; l >= 0; --l) {
    for (int q = P.lptr[l]; q < P.lptr[l + 1]; ++q) {
      const int f = P.lfront[q];
      if (in_sub[f]) continue;
      const int s2 = 2 * P.s[f], u2 = 2 * (P.sptr[f + 1] - P.sptr[f]);
      if (u2 == 0) continue;
      LevelTask t{};
      t.g0 = 2 * (int64_t)P.first[f]; t.s2 = s2; t.u2 = u2; t.f = f;
      // backward: one task per slab of the update set; a front with several slabs owns nk x s2p partial sums in `part`
      const int ks = bwd_ks(s2), nk = nbt[f], s2p = (s2 + 3) & ~3;
      t.goff = P.sptr[f]; t.nch = nk; t.pad[0] = ks;
      t.uoff = (int32_t)po;
      if (nk > 1) po += (int64_t)nk * s2p;
      if (po > INT32_MAX) throw StatusError(PLFEM_ERR_INVALID, "backward partial sums exceed the 32-bit offsets of the task records");
      const int par = P.parent[f];
      // exactly ONE signal per front (its single task, or whichever slab finishes): a child waits for 1; a root above has no
      // update set and no tasks — nothing to wait for
      t.dep = par; t.need = (par >= 0 && nbt[par] > 0) ? 1 : 0; t.sig = f;
      for (int k0 = 0; k0 < u2; k0 += ks) {
        t.r0 = k0; t.n = std::min(ks, u2 - k0);
        t.chunks = item_chunks(t.n, 32);                       // per warp
        t.soff = bo * CHD; bo += (int64_t)t.chunks * ((s2 + 31) >> 5);
        bt.push_back(t);
      }
    }
    S.bptr[l] = (int32_t)bt.size();           // level l: [bptr[l + 1], bptr[l])
  }
  S.bptr[P.nlevels] = 0;
  S.ftasks.upload(ctx, ft); S.btasks.upload(ctx, bt);
  S.lfwd.alloc(ctx, (size_t)std::max<int64_t>(fo, 1) * CHD); S.lbwd.alloc(ctx, (size_t)std::max<int64_t>(bo, 1) * CHD);
  S.lfwd_doubles = fo * CHD; S.lbwd_doubles = bo * CHD;
  S.sync.alloc(ctx, 2 + 3 * (size_t)P.nfronts);
  S.bpart.alloc(ctx, (size_t)std::max<int64_t>(po, 1) * SOLVE_NRHS);
  PLFEM_CUDA(stream_wait(ctx->stream));       // the host vectors above are pageable and local
}

void launch_stream_pack(plfem_ctx* ctx, const DevPlan& D) {
  const StreamPlan& S = D.st;
  if (S.n_fronts > 0) {
    stream_pack_kernel<<<S.n_fronts, 256, 0, ctx->stream>>>(S.recs.p, S.subs.p, D.s.p, D.sptr.p, D.foff.p, D.pool.p, S.lmaps.p, S.sfwd.p, S.sbwd.p);
    ctx->launches++;
  }
  const int nf = (int)S.ftasks.n, nb = (int)S.btasks.n;
  if (nf + nb > 0) {
    level_pack_kernel<<<nf + nb, 256, 0, ctx->stream>>>(S.ftasks.p, nf, S.btasks.p, D.foff.p, D.pool.p, S.lfwd.p, S.lbwd.p);
    ctx->launches++;
  }
  PLFEM_CUDA(cudaGetLastError());
}

namespace {
template <class... KArgs, class... Args>
void launch_warp_ctas(void (*kernel)(KArgs...), bool pdl, int grid, int threads, cudaStream_t st, Args... args) {
  static thread_local const void* configured[64] = {};
  bool seen = false;
  for (const void* k : configured) seen |= (k == (const void*)kernel);
  if (!seen) {      // all of the SM's shared memory for the rings: 8 warps per SM
    PLFEM_CUDA(cudaFuncSetAttribute((const void*)kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    for (const void*& k : configured) if (!k) { k = (const void*)kernel; break; }
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = 0; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
  PLFEM_CUDA(cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...));
}

// CTAs of a dataflow launch: PLFEM_SWEEP_CTAS_PER_SM (default 5, what the shared memory of the four rings allows) x SMs
int persistent_ctas(plfem_ctx* ctx) {
  static const int per_sm = [] { const char* e = std::getenv("PLFEM_SWEEP_CTAS_PER_SM"); return e ? std::max(1, std::min(16, atoi(e))) : 5; }();
  static thread_local int sms = 0;
  if (sms == 0) PLFEM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device));
  return per_sm * sms;
}

template <bool FUSED>
void launch_forward_tasks(plfem_ctx* ctx, const DevPlan& D, const LevelTask* tasks, int n, const double* rhs, double* out, int nrhs, bool pdl, const uint8_t* active) {
  const StreamPlan& S = D.st;
  const RhsView rv{rhs, out, D.upd.p};
  if (nrhs == 1) {
    if (pdl) launch_warp_ctas(level_forward_kernel<1, true, FUSED>, true, FUSED ? std::min(n, persistent_ctas(ctx)) : n, 32 * TW, ctx->stream, tasks, D.gsrc.p, S.lfwd.p, D.cptr.p, D.child.p, D.cmap_ptr.p, D.cmap.p, D.sptr.p, D.uoff.p, rv, D.status.p, S.sync.p, active, n);
    else launch_warp_ctas(level_forward_kernel<1, false, FUSED>, false, FUSED ? std::min(n, persistent_ctas(ctx)) : n, 32 * TW, ctx->stream, tasks, D.gsrc.p, S.lfwd.p, D.cptr.p, D.child.p, D.cmap_ptr.p, D.cmap.p, D.sptr.p, D.uoff.p, rv, D.status.p, S.sync.p, active, n);
  } else {
    if (pdl) launch_warp_ctas(level_forward_kernel<SOLVE_NRHS, true, FUSED>, true, FUSED ? std::min(n, persistent_ctas(ctx)) : n, 32 * TW, ctx->stream, tasks, D.gsrc.p, S.lfwd.p, D.cptr.p, D.child.p, D.cmap_ptr.p, D.cmap.p, D.sptr.p, D.uoff.p, rv, D.status.p, S.sync.p, active, n);
    else launch_warp_ctas(level_forward_kernel<SOLVE_NRHS, false, FUSED>, false, FUSED ? std::min(n, persistent_ctas(ctx)) : n, 32 * TW, ctx->stream, tasks, D.gsrc.p, S.lfwd.p, D.cptr.p, D.child.p, D.cmap_ptr.p, D.cmap.p, D.sptr.p, D.uoff.p, rv, D.status.p, S.sync.p, active, n);
  }
  ctx->launches++;
}

template <bool FUSED>
void launch_backward_tasks(plfem_ctx* ctx, const DevPlan& D, const LevelTask* tasks, int n, double* x, int nrhs, bool pdl, const uint8_t* active) {
  const StreamPlan& S = D.st;
  if (nrhs == 1) {
    if (pdl) launch_warp_ctas(level_backward_kernel<1, true, FUSED>, true, FUSED ? std::min(n, persistent_ctas(ctx)) : n, 32 * TW, ctx->stream, tasks, S.lbwd.p, D.strct.p, x, S.bpart.p, D.status.p, S.sync.p, S.nfronts, active, n);
    else launch_warp_ctas(level_backward_kernel<1, false, FUSED>, false, FUSED ? std::min(n, persistent_ctas(ctx)) : n, 32 * TW, ctx->stream, tasks, S.lbwd.p, D.strct.p, x, S.bpart.p, D.status.p, S.sync.p, S.nfronts, active, n);
  } else {
    if (pdl) launch_warp_ctas(level_backward_kernel<SOLVE_NRHS, true, FUSED>, true, FUSED ? std::min(n, persistent_ctas(ctx)) : n, 32 * TW, ctx->stream, tasks, S.lbwd.p, D.strct.p, x, S.bpart.p, D.status.p, S.sync.p, S.nfronts, active, n);
    else launch_warp_ctas(level_backward_kernel<SOLVE_NRHS, false, FUSED>, false, FUSED ? std::min(n, persistent_ctas(ctx)) : n, 32 * TW, ctx->stream, tasks, S.lbwd.p, D.strct.p, x, S.bpart.p, D.status.p, S.sync.p, S.nfronts, active, n);
  }
  ctx->launches++;
}
}  // namespace

void launch_stream_forward(plfem_ctx* ctx, const DevPlan& D, const double* rhs, double* out, int nrhs, bool pdl, const uint8_t* active) {
  const StreamPlan& S = D.st;
  if (S.n_subs == 0) return;
  if (nrhs == 1) {
    if (pdl) launch_warp_ctas(stream_forward_kernel<1, true>, true, S.n_subs, 32, ctx->stream, S.subs.p, S.fronts.p, S.sfwd.p, rhs, out, D.upd.p, D.status.p, active);
    else launch_warp_ctas(stream_forward_kernel<1, false>, false, S.n_subs, 32, ctx->stream, S.subs.p, S.fronts.p, S.sfwd.p, rhs, out, D.upd.p, D.status.p, active);
  } else {
    if (pdl) launch_warp_ctas(stream_forward_kernel<SOLVE_NRHS, true>, true, S.n_subs, 32, ctx->stream, S.subs.p, S.fronts.p, S.sfwd.p, rhs, out, D.upd.p, D.status.p, active);
    else launch_warp_ctas(stream_forward_kernel<SOLVE_NRHS, false>, false, S.n_subs, 32, ctx->stream, S.subs.p, S.fronts.p, S.sfwd.p, rhs, out, D.upd.p, D.status.p, active);
  }
  ctx->launches++;
}

void launch_level_forward(plfem_ctx* ctx, const DevPlan& D, int level, const double* rhs, double* out, int nrhs, bool pdl, const uint8_t* active) {
  const StreamPlan& S = D.st;
  launch_forward_tasks<false>(ctx, D, S.ftasks.p + S.fptr[level], S.fptr[level + 1] - S.fptr[level], rhs, out, nrhs, pdl, active);
}

void launch_level_backward(plfem_ctx* ctx, const DevPlan& D, int level, double* x, int nrhs, bool pdl, const uint8_t* active) {
  const StreamPlan& S = D.st;
  launch_backward_tasks<false>(ctx, D, S.btasks.p + S.bptr[level + 1], S.bptr[level] - S.bptr[level + 1], x, nrhs, pdl, active);
}

void set_sweep_trace(long long* p) { PLFEM_CUDA(cudaMemcpyToSymbol(g_sweep_trace, &p, sizeof(p))); }

void reset_sweep_counters(plfem_ctx* ctx, const DevPlan& D) {
  if (D.st.sync.n) PLFEM_CUDA(cudaMemsetAsync(D.st.sync.p, 0, D.st.sync.n * sizeof(int32_t), ctx->stream));
}

void launch_fused_forward(plfem_ctx* ctx, const DevPlan& D, const double* rhs, double* out, int nrhs, bool pdl, const uint8_t* active) {
  const StreamPlan& S = D.st;
  if (S.ftasks.n > 0) launch_forward_tasks<true>(ctx, D, S.ftasks.p, (int)S.ftasks.n, rhs, out, nrhs, pdl, active);
}

void launch_fused_backward(plfem_ctx* ctx, const DevPlan& D, double* x, int nrhs, bool pdl, const uint8_t* active) {
  const StreamPlan& S = D.st;
  if (S.btasks.n > 0) launch_backward_tasks<true>(ctx, D, S.btasks.p, (int)S.btasks.n, x, nrhs, pdl, active);
}

void launch_stream_backward(plfem_ctx* ctx, const DevPlan& D, double* x, int nrhs, bool pdl, const uint8_t* active) {
  const StreamPlan& S = D.st;
  if (S.n_subs == 0) return;
  if (nrhs == 1) {
    if (pdl) launch_warp_ctas(stream_backward_kernel<1, true>, true, S.n_subs, 32, ctx->stream, S.subs.p, S.fronts.p, S.sbwd.p, D.strct.p, x, D.status.p, active);
    else launch_warp_ctas(stream_backward_kernel<1, false>, false, S.n_subs, 32, ctx->stream, S.subs.p, S.fronts.p, S.sbwd.p, D.strct.p, x, D.status.p, active);
  } else {
    if (pdl) launch_warp_ctas(stream_backward_kernel<SOLVE_NRHS, true>, true, S.n_subs, 32, ctx->stream, S.subs.p, S.fronts.p, S.sbwd.p, D.strct.p, x, D.status.p, active);
    else launch_warp_ctas(stream_backward_kernel<SOLVE_NRHS, false>, false, S.n_subs, 32, ctx->stream, S.subs.p, S.fronts.p, S.sbwd.p, D.strct.p, x, D.status.p, active);
  }
  ctx->launches++;
}

}  // namespace plfem
