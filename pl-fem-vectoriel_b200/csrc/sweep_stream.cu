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
constexpr int NS = 4;                   // ring stages per warp
constexpr int NLOC = ST_NLOC;           // local unknowns (pivots of the subtree + update set of its root)
constexpr int MAXF = 32;                // fronts per subtree: one descriptor per lane
constexpr int RB = 128;                 // rows of a forward row block (4 rows per lane)

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

// The warp's view of its stream: a ring of NS chunks, refilled by lane 0 as soon as the warp has left a chunk.
struct Pipe {
  double* ring; uint64_t* bar; const double* src; int32_t* status;
  int nchunks, cur, pos, lane;
  bool dead = false;                                  // a wait timed out: stop waiting, the status flag reports it
  __device__ __forceinline__ void issue(int c) {
    const int s = c % NS;
    mbar_expect_tx(bar + s, CHD * 8);
    bulk_g2s(ring + s * CHD, src + (int64_t)c * CHD, CHD * 8, bar + s);
  }
  __device__ __forceinline__ void start() {
    if (lane == 0)
      for (int c = 0; c < NS && c < nchunks; ++c) issue(c);
    cur = -1; pos = CHD;
  }
  __device__ __forceinline__ void advance() {
    if (cur >= 0) {
      __syncwarp();                                   // every lane is done with the chunk being left
      if (lane == 0 && cur + NS < nchunks) issue(cur + NS);
    }
    ++cur; pos = 0;
    uint64_t* b = bar + cur % NS;
    const uint32_t parity = (cur / NS) & 1;
    unsigned spins = 0;
    while (!dead && !mbar_try_wait(b, parity))
      if (++spins > (1u << 18)) { atomicExch(status + 1, 2); dead = true; }   // report instead of hanging the GPU
  }
  __device__ __forceinline__ const double* place(int sz) {
    if (pos + sz > CHD) advance();
    const double* p = ring + (cur % NS) * CHD + pos;
    pos += sz;
    return p;
  }
  __device__ __forceinline__ void drain() {           // no copy may be in flight when the CTA exits
    while (cur + 1 < nchunks) advance();
  }
};

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

// forward, one row block (<= 128 rows, R per lane) of one front: acc = block * y1, y1 = loc[p0 ..) in place
template <int NR, int R>
__device__ __forceinline__ void fwd_block(Pipe& pp, StreamSmem<NR>& sm, int s2, int p0, int r0, int nrb, double* __restrict__ out,
                                          int64_t gpiv) {
  const int lane = pp.lane;
  const int ld = (nrb + 3) & ~3;
  bool valid[R];
#pragma unroll
  for (int q = 0; q < R; ++q) valid[q] = lane + 32 * q < nrb;
  double acc[R][NR];
#pragma unroll
  for (int q = 0; q < R; ++q)
#pragma unroll
    for (int r = 0; r < NR; ++r) acc[q][r] = 0.0;
  for (int k = 0; k < s2;) {
    int n = (CHD - pp.pos) / ld;
    if (n == 0) { pp.advance(); continue; }
    n = min(n, s2 - k);
    const double* base = pp.ring + (pp.cur % NS) * CHD + pp.pos + lane;
    const double* yv = sm.loc + (p0 + k) * NR;
#pragma unroll 4
    for (int c = 0; c < n; ++c) {
      double y[NR];
      ld_loc<NR>(yv, c, y);
#pragma unroll
      for (int q = 0; q < R; ++q) {
        const double m = valid[q] ? base[c * ld + 32 * q] : 0.0;
#pragma unroll
        for (int r = 0; r < NR; ++r) acc[q][r] = fma(m, y[r], acc[q][r]);
      }
    }
    pp.pos += n * ld; k += n;
  }
#pragma unroll
  for (int q = 0; q < R; ++q) {
    if (!valid[q]) continue;
    const int row = r0 + lane + 32 * q;
    if (row < s2) {
      stg_v<NR>(out, gpiv + row, acc[q]);                        // z1 = F11^-1 y1
    } else {
      const int j = row - s2;
      double* t = sm.loc + (2 * (int)sm.lm[j >> 1] + (j & 1)) * NR;   // the ancestor's local position: y -= W^T y1
#pragma unroll
      for (int r = 0; r < NR; ++r) t[r] -= acc[q][r];
    }
  }
}

template <int NR, bool PDL>
__global__ void __launch_bounds__(32) stream_forward_kernel(const StreamSub* __restrict__ subs, const int4* __restrict__ fronts,
                                                            const double* __restrict__ stream, const double* __restrict__ rhs,
                                                            double* __restrict__ out, double* __restrict__ upd, int32_t* status) {
  __shared__ StreamSmem<NR> sm;
  const int lane = threadIdx.x;
  if (PDL) griddep_launch_dependents();
  const StreamSub sb = subs[blockIdx.x];
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
  for (int i = lane; i < sb.nI + sb.next; i += 32) {
    double v[NR];
    if (i < sb.nI) {
      ldg_v<NR>(rhs, (int64_t)sb.g0 + i, v);
    } else {
#pragma unroll
      for (int r = 0; r < NR; ++r) v[r] = 0.0;
    }
#pragma unroll
    for (int r = 0; r < NR; ++r) sm.loc[i * NR + r] = v[r];
  }
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
  double acc[R][NR];
#pragma unroll
  for (int q = 0; q < R; ++q)
#pragma unroll
    for (int r = 0; r < NR; ++r) acc[q][r] = 0.0;
  for (int j = 0; j < u2;) {
    int n = (CHD - pp.pos) / s2p;
    if (n == 0) { pp.advance(); continue; }
    n = min(n, u2 - j);
    const double* base = pp.ring + (pp.cur % NS) * CHD + pp.pos + lane;
#pragma unroll 4
    for (int c = 0; c < n; ++c) {
      const int jj = j + c;
      double x[NR];
      ld_loc<NR>(sm.loc, 2 * (int)sm.lm[jj >> 1] + (jj & 1), x);
#pragma unroll
      for (int q = 0; q < R; ++q) {
        const double m = valid[q] ? base[c * s2p + 32 * q] : 0.0;
#pragma unroll
        for (int r = 0; r < NR; ++r) acc[q][r] = fma(m, x[r], acc[q][r]);
      }
    }
    pp.pos += n * s2p; j += n;
  }
#pragma unroll
  for (int q = 0; q < R; ++q) {
    if (!valid[q]) continue;
    double* t = sm.loc + (p0 + lane + 32 * q) * NR;
#pragma unroll
    for (int r = 0; r < NR; ++r) t[r] -= acc[q][r];
  }
}

template <int NR, bool PDL>
__global__ void __launch_bounds__(32) stream_backward_kernel(const StreamSub* __restrict__ subs, const int4* __restrict__ fronts,
                                                             const double* __restrict__ stream, const int32_t* __restrict__ strct,
                                                             double* __restrict__ x, int32_t* status) {
  __shared__ StreamSmem<NR> sm;
  const int lane = threadIdx.x;
  if (PDL) griddep_launch_dependents();
  const StreamSub sb = subs[blockIdx.x];
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
  for (int i = lane; i < sb.nI; i += 32) {                        // z of the subtree's pivots (forward sweep)
    double v[NR];
    ldg_v<NR>(x, (int64_t)sb.g0 + i, v);
#pragma unroll
    for (int r = 0; r < NR; ++r) sm.loc[i * NR + r] = v[r];
  }
#pragma unroll
  for (int t = 0; t < NE; ++t) {                                  // x of the ancestors above the subtree (final)
    if (xo[t] < 0) continue;
    double v[NR];
    ldg_v<NR>(x, xo[t], v);
    const int e = lane + 32 * t;
#pragma unroll
    for (int r = 0; r < NR; ++r) sm.loc[(sb.nI + e) * NR + r] = v[r];
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

void launch_stream_pack(plfem_ctx* ctx, const DevPlan& D) {
  const StreamPlan& S = D.st;
  if (S.n_fronts == 0) return;
  stream_pack_kernel<<<S.n_fronts, 256, 0, ctx->stream>>>(S.recs.p, S.subs.p, D.s.p, D.sptr.p, D.foff.p, D.pool.p, S.lmaps.p, S.sfwd.p, S.sbwd.p);
  PLFEM_CUDA(cudaGetLastError());
  ctx->launches++;
}

namespace {
template <class... KArgs, class... Args>
void launch_warp_ctas(void (*kernel)(KArgs...), bool pdl, int grid, cudaStream_t st, Args... args) {
  static thread_local const void* configured[8] = {};
  bool seen = false;
  for (const void* k : configured) seen |= (k == (const void*)kernel);
  if (!seen) {      // all of the SM's shared memory for the rings: 8 warps per SM
    PLFEM_CUDA(cudaFuncSetAttribute((const void*)kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    for (const void*& k : configured) if (!k) { k = (const void*)kernel; break; }
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(32); cfg.dynamicSmemBytes = 0; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
  PLFEM_CUDA(cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...));
}
}  // namespace

void launch_stream_forward(plfem_ctx* ctx, const DevPlan& D, const double* rhs, double* out, int nrhs, bool pdl) {
  const StreamPlan& S = D.st;
  if (S.n_subs == 0) return;
  if (nrhs == 1) {
    if (pdl) launch_warp_ctas(stream_forward_kernel<1, true>, true, S.n_subs, ctx->stream, S.subs.p, S.fronts.p, S.sfwd.p, rhs, out, D.upd.p, D.status.p);
    else launch_warp_ctas(stream_forward_kernel<1, false>, false, S.n_subs, ctx->stream, S.subs.p, S.fronts.p, S.sfwd.p, rhs, out, D.upd.p, D.status.p);
  } else {
    if (pdl) launch_warp_ctas(stream_forward_kernel<SOLVE_NRHS, true>, true, S.n_subs, ctx->stream, S.subs.p, S.fronts.p, S.sfwd.p, rhs, out, D.upd.p, D.status.p);
    else launch_warp_ctas(stream_forward_kernel<SOLVE_NRHS, false>, false, S.n_subs, ctx->stream, S.subs.p, S.fronts.p, S.sfwd.p, rhs, out, D.upd.p, D.status.p);
  }
  ctx->launches++;
}

void launch_stream_backward(plfem_ctx* ctx, const DevPlan& D, double* x, int nrhs, bool pdl) {
  const StreamPlan& S = D.st;
  if (S.n_subs == 0) return;
  if (nrhs == 1) {
    if (pdl) launch_warp_ctas(stream_backward_kernel<1, true>, true, S.n_subs, ctx->stream, S.subs.p, S.fronts.p, S.sbwd.p, D.strct.p, x, D.status.p);
    else launch_warp_ctas(stream_backward_kernel<1, false>, false, S.n_subs, ctx->stream, S.subs.p, S.fronts.p, S.sbwd.p, D.strct.p, x, D.status.p);
  } else {
    if (pdl) launch_warp_ctas(stream_backward_kernel<SOLVE_NRHS, true>, true, S.n_subs, ctx->stream, S.subs.p, S.fronts.p, S.sbwd.p, D.strct.p, x, D.status.p);
    else launch_warp_ctas(stream_backward_kernel<SOLVE_NRHS, false>, false, S.n_subs, ctx->stream, S.subs.p, S.fronts.p, S.sbwd.p, D.strct.p, x, D.status.p);
  }
  ctx->launches++;
}

}  // namespace plfem
