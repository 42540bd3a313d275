// K4-K6: level-batched multifrontal block-LDL^T factorisation of (A - sigma B) and the two sweeps.
//
// Replaces SuperLU `splu` + `solve` inside scipy eigsh (solver_fem.py:197).
//
// Design (B200): the shifted operator is symmetric indefinite and only ~45k-2M unknowns, so the
// work is latency-bound, not flop-bound.  Fronts come from a host nested dissection (symbolic.cpp);
// every level of the elimination tree is ONE batched launch per stage, all fronts of the level in
// flight at once:
//     extend-add  ->  invert pivot block  ->  W^T = (F11^-1 F12)^T  ->  S = F22 - F12^T W
// A front is a dense column-major (2nf x 2nf) matrix (two unknowns, Hx and Hy, per P2 node).  The
// pivot block F11 (<= 128 x 128) is inverted explicitly with partial pivoting in shared memory, so
// the solve phase is pure matrix-vector work on the contiguous left block column [F11^-1 ; W^T]:
//     forward :  z1 = F11^-1 y1 ,  upd = y2 - W^T y1      (one coalesced GEMV over 2nf rows)
//     backward:  x1 = z1 - W x2                            (one dot product per column of W^T)
// Contributions travel child -> parent through per-front update blocks/vectors with precomputed
// position maps: gather-only, fixed order, no atomics, bit-reproducible run to run.
#include "common.h"

#include <cooperative_groups.h>

#include <algorithm>
#include <cstdlib>

namespace plfem {

namespace {

constexpr int GT = 64;        // GEMM tile (GT x GT outputs per CTA)
constexpr int GK = 16;        // GEMM k-step
constexpr int EA_COLS = 8;    // extend-add slab width in parent node columns
constexpr int MAX_PIV = 128;  // largest pivot block (unknowns) the in-shared-memory inverse handles

struct PlanView {
  const int32_t *first, *s, *sptr, *strct, *cptr, *child, *cmap_ptr, *cmap, *uoff;
  const int64_t* foff;
  double* pool;
};

__device__ __forceinline__ int front_u(const PlanView& P, int f) { return P.sptr[f + 1] - P.sptr[f]; }

// ---- load (A - sigma B) into the fronts: one thread per structural non-zero ---------------------------
__global__ void front_load_kernel(int64_t nnz, const int32_t* __restrict__ rowidx, const int32_t* __restrict__ col,
                                  const int32_t* __restrict__ sn_of, PlanView P, const double* __restrict__ vals,
                                  const double* __restrict__ sigma_node) {
  const int64_t z = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (z >= nnz) return;
  const int32_t r = rowidx[z], c = col[z];
  const int32_t f = sn_of[r];
  const int32_t f0 = P.first[f], s = P.s[f];
  if (c < f0) return;  // belongs to the front that eliminates c
  int32_t pc;
  if (c < f0 + s) {
    pc = c - f0;
  } else {
    int32_t lo = P.sptr[f], hi = P.sptr[f + 1];
    const int32_t base = lo;
    while (lo < hi) {
      const int32_t mid = (lo + hi) >> 1;
      if (P.strct[mid] < c) lo = mid + 1; else hi = mid;
    }
    pc = s + (lo - base);
  }
  const int64_t ld = 2 * (int64_t)(s + front_u(P, f));
  double* F = P.pool + P.foff[f];
  const int64_t pr = r - f0;
  const double b = sigma_node[r] * vals[(int64_t)S_MINV * nnz + z];   // one shift per design of the forest
  F[(2 * pc) * ld + 2 * pr] = vals[(int64_t)S_AXX * nnz + z] - b;
  F[(2 * pc + 1) * ld + 2 * pr] = vals[(int64_t)S_AXY * nnz + z];
  F[(2 * pc) * ld + 2 * pr + 1] = vals[(int64_t)S_AYX * nnz + z];
  F[(2 * pc + 1) * ld + 2 * pr + 1] = vals[(int64_t)S_AYY * nnz + z] - b;
}

// ---- extend-add: parent += children's Schur complements, one CTA per (parent, column slab) --------------
__global__ void __launch_bounds__(256) extend_add_kernel(const int4* __restrict__ slabs, PlanView P) {
  const int4 sl = slabs[blockIdx.x];
  const int f = sl.x, c0 = sl.y, c1 = sl.z;
  const int sp = P.s[f];
  const int64_t ldp = 2 * (int64_t)(sp + front_u(P, f));
  double* Fp = P.pool + P.foff[f];
  for (int q = P.cptr[f]; q < P.cptr[f + 1]; ++q) {
    const int ch = P.child[q];
    const int sc = P.s[ch], uc = front_u(P, ch);
    const int32_t* cm = P.cmap + P.cmap_ptr[ch];
    // child columns whose parent position falls in [c0, c1)  (cm is increasing)
    int lo = 0, hi = uc;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (cm[mid] < c0) lo = mid + 1; else hi = mid; }
    const int jlo = lo;
    hi = uc;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (cm[mid] < c1) lo = mid + 1; else hi = mid; }
    const int jhi = lo;
    const int64_t ldc = 2 * (int64_t)(sc + uc);
    const double* Sc = P.pool + P.foff[ch] + (2 * (int64_t)sc) * ldc + 2 * sc;  // F22 of the child
    const int rows = 2 * uc, cols = 2 * (jhi - jlo);
    for (int64_t t = threadIdx.x; t < (int64_t)rows * cols; t += blockDim.x) {
      const int i = (int)(t % rows), j = (int)(t / rows) + 2 * jlo;
      const int pj = cm[j >> 1], pi = cm[i >> 1];
      if (pi >= sp && pj < sp) continue;  // F21 of the parent is never read
      Fp[(2 * (int64_t)pj + (j & 1)) * ldp + 2 * pi + (i & 1)] += Sc[(int64_t)j * ldc + i];
    }
    __syncthreads();  // two children may hit the same parent entry: keep child order
  }
}

// ---- pivot block inverse: Gauss-Jordan with partial pivoting in shared memory, one CTA per front ------
// Thread layout: i = tid % MP (row), jg = tid / MP (column group), MP = m rounded up to 32/64/128, so the
// rank-1 update of step k needs no integer division and touches shared memory conflict-free.
__global__ void __launch_bounds__(1024) invert_kernel(const int32_t* __restrict__ fronts, PlanView P, int32_t* status) {
  extern __shared__ double sm[];
  const int f = fronts[blockIdx.x];
  const int m = 2 * P.s[f];
  const int64_t ld = 2 * (int64_t)(P.s[f] + front_u(P, f));
  double* F = P.pool + P.foff[f];
  const int lds = m | 1;
  double* a = sm;                 // m x m, column-major, leading dimension lds
  double* colk = a + (size_t)lds * m;
  double* rowk = colk + m;
  __shared__ int piv[MAX_PIV];
  __shared__ int src[MAX_PIV];
  __shared__ int s_p;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int MP = (m <= 32) ? 32 : (m <= 64 ? 64 : 128);
  const int i = tid & (MP - 1), jg = tid / MP, ng = nt / MP;
  if (i < m)
    for (int j = jg; j < m; j += ng) a[i + j * lds] = F[(int64_t)j * ld + i];
  __syncthreads();
  for (int k = 0; k < m; ++k) {
    if (tid < 32) {
      double best = -1.0; int bi = k;
      for (int r = k + tid; r < m; r += 32) {
        const double v = fabs(a[r + k * lds]);
        if (v > best) { best = v; bi = r; }   // NaN never wins; an all-NaN column keeps bi = k
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        const double ob = __shfl_down_sync(0xffffffffu, best, off);
        const int oi = __shfl_down_sync(0xffffffffu, bi, off);
        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
      }
      if (tid == 0) {
        s_p = bi; piv[k] = bi;
        if (!(best > 0.0) || !isfinite(best)) { atomicExch(status, 1); atomicExch(status + 3, f); }
      }
    }
    __syncthreads();
    const int p = s_p;
    const double inv = 1.0 / a[p + k * lds];
    // row swap k <-> p fused with the copies of the scaled pivot row (first m threads) and of the
    // pivot column (last m threads)
    if (tid < m) {
      const int j = tid;
      const double ak = a[k + j * lds], ap = a[p + j * lds];
      rowk[j] = ap * inv;
      if (p != k && j != k) a[p + j * lds] = ak;
    }
    if (tid >= nt - m) {
      const int r = tid - (nt - m);
      double v;
      if (r == k) v = 0.0;                         // unused
      else if (r == p) v = a[k + k * lds];         // row p now holds old row k
      else v = a[r + k * lds];
      colk[r] = v;
    }
    __syncthreads();
    if (i < m) {
      const double ci = colk[i];
      if (i == k) {
        for (int j = jg; j < m; j += ng) a[i + j * lds] = (j == k) ? inv : rowk[j];
      } else {
        for (int j = jg; j < m; j += ng) a[i + j * lds] = (j == k) ? -ci * inv : a[i + j * lds] - ci * rowk[j];
      }
    }
    __syncthreads();
  }
  if (tid == 0) {
    for (int j = 0; j < m; ++j) src[j] = j;
    for (int k = m - 1; k >= 0; --k) { const int t = src[k]; src[k] = src[piv[k]]; src[piv[k]] = t; }
  }
  __syncthreads();
  // The inverse of a symmetric block is symmetric; the computed one only to cond x eps, and the Schur complement
  // S = F22 - F12^T (F11^-1 F12) is formed from the upper block alone, so that antisymmetric part would land in S, then in
  // the parent's pivot block, amplified by |W|^2 level by level (it, not the block-local pivoting, cost
  // the raw solve six digits and made structured meshes diverge).  Write back the average with the transpose.
  if (i < m)
    for (int j = jg; j < m; j += ng) F[(int64_t)j * ld + i] = 0.5 * (a[i + src[j] * lds] + a[j + src[i] * lds]);
}

// ---- tiled FP64 GEMM used for W^T and the Schur update -------------------------------------------------
// C(i,j) (+)= sign * sum_k A(i,k) * B(k,j); A is k-contiguous (A(i,k) = Ap[i*lda + k]);
// B is either k-contiguous (B(k,j) = Bp[j*ldb + k]) or j-contiguous (Bp[k*ldb + j]); C(i,j) = Cp[j*ldc + i].
template <bool B_KCONTIG, bool ACCUM>
__device__ __forceinline__ void gemm_tile(const double* __restrict__ Ap, int64_t lda, const double* __restrict__ Bp,
                                          int64_t ldb, double* __restrict__ Cp, int64_t ldc, int M, int N, int K,
                                          int m0, int n0) {
  __shared__ double As[GK][GT + 1];
  __shared__ double Bs[GK][GT + 1];
  const int tid = threadIdx.x;          // 256 threads, 16 x 16, each 4 x 4 outputs
  const int tx = tid % 16, ty = tid / 16;
  double acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
  for (int k0 = 0; k0 < K; k0 += GK) {
    // A tile: GT rows x GK k, k fastest in memory
    for (int t = tid; t < GT * GK; t += 256) {
      const int kk = t % GK, i = t / GK;
      const int gi = m0 + i, gk = k0 + kk;
      As[kk][i] = (gi < M && gk < K) ? Ap[(int64_t)gi * lda + gk] : 0.0;
    }
    if (B_KCONTIG) {
      for (int t = tid; t < GT * GK; t += 256) {
        const int kk = t % GK, j = t / GK;
        const int gj = n0 + j, gk = k0 + kk;
        Bs[kk][j] = (gj < N && gk < K) ? Bp[(int64_t)gj * ldb + gk] : 0.0;
      }
    } else {
      for (int t = tid; t < GT * GK; t += 256) {
        const int j = t % GT, kk = t / GT;
        const int gj = n0 + j, gk = k0 + kk;
        Bs[kk][j] = (gj < N && gk < K) ? Bp[(int64_t)gk * ldb + gj] : 0.0;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < GK; ++kk) {
      double av[4], bv[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) av[a] = As[kk][tx + 16 * a];
#pragma unroll
      for (int b = 0; b < 4; ++b) bv[b] = Bs[kk][ty + 16 * b];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = fma(av[a], bv[b], acc[a][b]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int b = 0; b < 4; ++b) {
    const int gj = n0 + ty + 16 * b;
    if (gj >= N) continue;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int gi = m0 + tx + 16 * a;
      if (gi >= M) continue;
      double* c = Cp + (int64_t)gj * ldc + gi;
      if (ACCUM) *c -= acc[a][b]; else *c = acc[a][b];
    }
  }
}

// W^T:  F21new(j, i) = sum_k F12(k, j) * F11inv(k, i)      (j over 2u, i over 2s, k over 2s)
__global__ void __launch_bounds__(256) gemm_w_kernel(const int4* __restrict__ tiles, PlanView P) {
  const int4 t = tiles[blockIdx.x];
  const int f = t.x;
  const int s2 = 2 * P.s[f], u2 = 2 * front_u(P, f);
  const int64_t ld = s2 + u2;
  double* F = P.pool + P.foff[f];
  gemm_tile<true, false>(F + (int64_t)s2 * ld /* F12: A(j,k) at (s2+j)*ld + k */, ld, F /* F11inv(k,i) at i*ld + k */, ld,
                         F + s2 /* F21(j,i) at i*ld + s2 + j */, ld, u2, s2, s2, t.y, t.z);
}

// Schur:  F22(a, b) -= sum_k F12(k, a) * F21new(b, k)        (a, b over 2u, k over 2s)
__global__ void __launch_bounds__(256) gemm_schur_kernel(const int4* __restrict__ tiles, PlanView P) {
  const int4 t = tiles[blockIdx.x];
  const int f = t.x;
  const int s2 = 2 * P.s[f], u2 = 2 * front_u(P, f);
  const int64_t ld = s2 + u2;
  double* F = P.pool + P.foff[f];
  gemm_tile<false, true>(F + (int64_t)s2 * ld, ld, F + s2 /* B(k,b) = F21new(b,k) at k*ld + s2 + b */, ld,
                         F + (int64_t)s2 * ld + s2, ld, u2, u2, s2, t.y, t.z);
}

// ---- sweep work items --------------------------------------------------------------------------------------
// A sweep is a chain of ~2 x levels dependent steps, each a handful of microseconds of which most used to be
// dependent L2 round trips chasing plan metadata (slab -> front -> children -> maps -> values).  Every item
// now carries all it needs in one 96-byte record, the child->parent maps are flattened per parent row into
// two source offsets (first and second child), and everything static — the record, the source offsets, the
// factor entries the thread will multiply — is loaded BEFORE the item waits for its dependencies, so the
// critical path of a step is: see the flag, one gather of the freshly written values, FMAs from registers,
// write, signal.
template <int NR>
struct SweepSmem {
  double y[MAX_PIV][NR];          // assembled pivot part of the right-hand sides, [k][rhs] (one 16/32-byte read per k)
  double part[8][2][NR][32];      // partial sums: [warp][row of the thread][rhs][lane]
};

// The NR right-hand sides of a sweep are INTERLEAVED: entry i of all of them is rhs[i*NR .. i*NR+NR), one 32-byte
// sector for NR = 4.  Every gather of the sweeps (children's updates, ancestors' unknowns) fetches all right-hand
// sides of an index at once, so this costs one sector and two 128-bit loads per index instead of four of each.
struct RhsView {
  const double* rhs; double* out; double* upd;
};

// CG = true: values produced by other CTAs of the SAME launch (dataflow kernel) are read with ld.global.cg,
// i.e. from L2, never from a possibly stale L1 line.
template <bool CG>
__device__ __forceinline__ double ldx(const double* p) { return CG ? __ldcg(p) : *p; }

// all NR right-hand sides of entry idx of an interleaved vector block
template <bool CG, int NR>
__device__ __forceinline__ void ldv(const double* base, int64_t idx, double (&v)[NR]) {
  if constexpr (NR % 2 == 0) {
    const double2* p = reinterpret_cast<const double2*>(base + idx * NR);
#pragma unroll
    for (int r = 0; r < NR / 2; ++r) { const double2 t = CG ? __ldcg(p + r) : p[r]; v[2 * r] = t.x; v[2 * r + 1] = t.y; }
  } else {
#pragma unroll
    for (int r = 0; r < NR; ++r) v[r] = ldx<CG>(base + idx * NR + r);
  }
}
template <int NR>
__device__ __forceinline__ void stv(double* base, int64_t idx, const double (&v)[NR]) {
  if constexpr (NR % 2 == 0) {
    double2* p = reinterpret_cast<double2*>(base + idx * NR);
#pragma unroll
    for (int r = 0; r < NR / 2; ++r) p[r] = make_double2(v[2 * r], v[2 * r + 1]);
  } else {
#pragma unroll
    for (int r = 0; r < NR; ++r) base[idx * NR + r] = v[r];
  }
}

__device__ __forceinline__ void wait_count(const int32_t* ctr, int32_t target, int32_t* status) {
  const volatile int32_t* v = ctr;
  unsigned spins = 0;
  while (*v < target) {
    if (++spins > (1u << 22)) { atomicExch(status + 1, 1); break; }   // ~seconds: report instead of hanging
  }
}

// Programmatic dependent launch: a kernel launched with the programmatic-serialisation attribute may start
// while its predecessor in the stream is still running; everything before griddep_wait() (the static
// prefetch) overlaps the predecessor's tail, everything after it sees the predecessor's writes.
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

struct Deps {          // dataflow bookkeeping of the persistent kernel (unused by the per-level kernels)
  int32_t* fdone; int32_t* bdone; int32_t* status; int epoch; const int32_t* nfs;
};

constexpr int FQ = 16;  // factor entries per thread in flight before the wait

// NR doubles of one row of a [.][NR] shared array: a single 128-bit load per pair of right-hand sides
template <int NR>
__device__ __forceinline__ void lds_row(const double (*a)[NR], int k, double (&v)[NR]) {
  if constexpr (NR % 2 == 0) {
    const double2* p = reinterpret_cast<const double2*>(a[k]);
#pragma unroll
    for (int r = 0; r < NR / 2; ++r) { const double2 t = p[r]; v[2 * r] = t.x; v[2 * r + 1] = t.y; }
  } else {
#pragma unroll
    for (int r = 0; r < NR; ++r) v[r] = a[k][r];
  }
}

// forward: one CTA (8 warps) per (front, slab of 32*G*R rows) of the packed left block column.  The 8 warps form G
// row groups x 8/G slices of the k range (the 2s pivot columns), every thread owns R rows 32 apart; partial sums
// meet in shared memory.  The loops are lean on purpose — ncu showed the previous version issue-bound at ~80 warp
// instructions per loaded factor entry: loop bounds are warp-uniform, rows past the end of the slab read row 0
// (and are discarded) instead of predicating every load, the assembled right-hand sides are read with one 128-bit
// shared load per pair, pointers advance by a constant stride.
template <bool CG, int NR, int R, bool PDL>
__device__ __forceinline__ void forward_item_r(const FwdItem& it, const int32_t* __restrict__ gsrc, const PlanView& P,
                                               const RhsView& rv, SweepSmem<NR>& sm, const Deps& dp) {
  constexpr int FQR = FQ / R;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int s2 = it.s2, nrows = it.nrows, row0 = it.row0, G = it.G & 0xff, nf2 = it.nf2;
  const int64_t ld = it.ld;
  const int nks = 8 / G, rg = warp % G, ks = warp / G;
  // ---- static prefetch: gather sources and this thread's factor entries
  const int32_t* g1 = gsrc + it.goff;       // first child: source offset into upd per front row, -1 = none
  const int32_t* g2 = g1 + nf2;             // second child
  int i1 = -1, i2 = -1;
  if (tid < s2) { i1 = g1[tid]; i2 = g2[tid]; }
  int lr[R]; bool ok[R]; int j1[R], j2[R];
#pragma unroll
  for (int q = 0; q < R; ++q) {
    lr[q] = (rg * R + q) * 32 + lane;
    ok[q] = lr[q] < nrows;
    j1[q] = j2[q] = -1;
    if (ks == 0 && ok[q] && row0 + lr[q] >= s2) { j1[q] = g1[row0 + lr[q]]; j2[q] = g2[row0 + lr[q]]; }
  }
  const int nk = (s2 - ks + nks - 1) / nks;          // k's of this warp's slice: ks, ks + nks, ...  (warp-uniform)
  const int64_t step = (int64_t)nks * ld;
  const double* p[R];
#pragma unroll
  for (int q = 0; q < R; ++q) p[q] = P.pool + it.foff + row0 + (ok[q] ? lr[q] : 0) + (int64_t)ks * ld;
  double m[FQR][R];
#pragma unroll
  for (int i = 0; i < FQR; ++i) {
    if (i < nk) {
#pragma unroll
      for (int q = 0; q < R; ++q) m[i][q] = p[q][i * step];
    } else {
#pragma unroll
      for (int q = 0; q < R; ++q) m[i][q] = 0.0;
    }
  }
  if (PDL) griddep_wait();                 // the previous level's kernel is complete and visible from here on
  // ---- wait for the children (dataflow mode)
  if (CG) {
    if (tid == 0) {
      if (it.ch0 >= 0) wait_count(dp.fdone + it.ch0, it.tgt0 * dp.epoch, dp.status);
      if (it.ch1 >= 0) wait_count(dp.fdone + it.ch1, it.tgt1 * dp.epoch, dp.status);
      for (int q = 2; q < it.nchild; ++q) { const int ch = P.child[P.cptr[it.f] + q]; wait_count(dp.fdone + ch, dp.nfs[ch] * dp.epoch, dp.status); }
    }
    __syncthreads();
  }
  // ---- dynamic part: right-hand side + children's updates, fixed order (rhs + first child) + second child
  if (tid < s2) {
    double v[NR], w[NR];
    ldv<CG, NR>(rv.rhs, it.g0 + tid, v);
    if (i1 >= 0) {
      ldv<CG, NR>(rv.upd, i1, w);
#pragma unroll
      for (int r = 0; r < NR; ++r) v[r] += w[r];
    }
    if (i2 >= 0) {
      ldv<CG, NR>(rv.upd, i2, w);
#pragma unroll
      for (int r = 0; r < NR; ++r) v[r] += w[r];
    }
#pragma unroll
    for (int r = 0; r < NR; ++r) sm.y[tid][r] = v[r];
  }
  double yt[R][NR];                          // what the children send to this thread's update rows (final-stage threads)
#pragma unroll
  for (int q = 0; q < R; ++q) {
    double w[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) yt[q][r] = 0.0;
    if (j1[q] >= 0) {
      ldv<CG, NR>(rv.upd, j1[q], w);
#pragma unroll
      for (int r = 0; r < NR; ++r) yt[q][r] += w[r];
    }
    if (j2[q] >= 0) {
      ldv<CG, NR>(rv.upd, j2[q], w);
#pragma unroll
      for (int r = 0; r < NR; ++r) yt[q][r] += w[r];
    }
  }
  __syncthreads();
  if (it.nchild > 2) {                       // rare (a separator that does not disconnect): generic path
    for (int c = P.cptr[it.f] + 2; c < P.cptr[it.f + 1]; ++c) {
      const int ch = P.child[c];
      const int uc2 = 2 * front_u(P, ch);
      const int32_t* cm = P.cmap + P.cmap_ptr[ch];
      for (int r = 0; r < NR; ++r) {
        const double* uv = rv.upd + (int64_t)P.uoff[ch] * NR + r;      // entry k of right-hand side r at uv[k * NR]
        for (int k = tid; k < uc2; k += 256) {
          const int t = 2 * cm[k >> 1] + (k & 1);
          if (t < s2) sm.y[t][r] += ldx<CG>(uv + (int64_t)k * NR);
        }
        if (ks == 0) {
          for (int k = 0; k < uc2; ++k) {
            const int t = 2 * cm[k >> 1] + (k & 1);
#pragma unroll
            for (int q = 0; q < R; ++q) if (ok[q] && t == row0 + lr[q] && t >= s2) yt[q][r] += ldx<CG>(uv + (int64_t)k * NR);
          }
        }
      }
      __syncthreads();
    }
  }
  double acc[R][NR];
#pragma unroll
  for (int q = 0; q < R; ++q)
#pragma unroll
    for (int r = 0; r < NR; ++r) acc[q][r] = 0.0;
#pragma unroll
  for (int i = 0; i < FQR; ++i) {
    if (i < nk) {
      double y[NR];
      lds_row<NR>(sm.y, ks + i * nks, y);
#pragma unroll
      for (int q = 0; q < R; ++q)
#pragma unroll
        for (int r = 0; r < NR; ++r) acc[q][r] = fma(m[i][q], y[r], acc[q][r]);
    }
  }
#pragma unroll 4
  for (int i = FQR; i < nk; ++i) {
    double mv[R], y[NR];
#pragma unroll
    for (int q = 0; q < R; ++q) mv[q] = p[q][i * step];
    lds_row<NR>(sm.y, ks + i * nks, y);
#pragma unroll
    for (int q = 0; q < R; ++q)
#pragma unroll
      for (int r = 0; r < NR; ++r) acc[q][r] = fma(mv[q], y[r], acc[q][r]);
  }
#pragma unroll
  for (int q = 0; q < R; ++q)
#pragma unroll
    for (int r = 0; r < NR; ++r) sm.part[warp][q][r][lane] = acc[q][r];
  __syncthreads();
  if (ks == 0) {
#pragma unroll
    for (int q = 0; q < R; ++q) {
      if (!ok[q]) continue;
      const int row = row0 + lr[q];
      double o[NR];
#pragma unroll
      for (int r = 0; r < NR; ++r) {
        double t = 0.0;
        for (int c = 0; c < nks; ++c) t += sm.part[c * G + rg][q][r][lane];
        o[r] = row < s2 ? t : yt[q][r] - t;
      }
      if (row < s2) stv<NR>(rv.out, it.g0 + row, o);
      else stv<NR>(rv.upd, (int64_t)it.uoff + (row - s2), o);
    }
  }
}

template <bool CG, int NR, bool PDL = false>
__device__ __forceinline__ void forward_item(const FwdItem& it, const int32_t* __restrict__ gsrc, const PlanView& P,
                                             const RhsView& rv, SweepSmem<NR>& sm, const Deps& dp) {
  if ((it.G >> 8) == 2) forward_item_r<CG, NR, 2, PDL>(it, gsrc, P, rv, sm, dp);
  else forward_item_r<CG, NR, 1, PDL>(it, gsrc, P, rv, sm, dp);
}

// backward: one CTA per (front, chunk of 32*R pivot columns) against the row-major copy of W (lanes over pivot
// columns, coalesced rows).  The update unknowns x2 are gathered ONCE per CTA into shared memory — the gather costs
// as many loads as the factor chunk itself, so it must not be repeated per column — and the 8 warps take slices of
// the j range (the 2u update unknowns); partial sums meet in shared memory.  Same lean loops as the forward item.
constexpr int JT = 512;    // update unknowns staged per pass (one pass for all fronts met so far)
template <int NR>
struct BwdSmem {
  double xs[JT][NR];
  double part[8][2][NR][32];
};

template <bool CG, int NR, int R, bool PDL>
__device__ __forceinline__ void backward_item_r(const BwdItem& it, const PlanView& P, double* x, BwdSmem<NR>& sm,
                                                const Deps& dp) {
  constexpr int FQR = FQ / R;
  constexpr int nks = 8;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int u2 = it.u2, ks = warp;
  const int64_t s2p = it.ld;
  const int32_t* st = P.strct + it.soff;
  int lc[R]; bool ok[R];
#pragma unroll
  for (int q = 0; q < R; ++q) { lc[q] = q * 32 + lane; ok[q] = lc[q] < it.ncols; }
  // ---- static prefetch: gather offsets of the first pass and this thread's factor entries
  int64_t o0 = -1, o1 = -1;
  if (tid < u2) o0 = 2 * (int64_t)st[tid >> 1] + (tid & 1);
  if (tid + 256 < u2) o1 = 2 * (int64_t)st[(tid + 256) >> 1] + (tid & 1);
  const double* p[R];
#pragma unroll
  for (int q = 0; q < R; ++q) p[q] = P.pool + it.foff + it.col0 + (ok[q] ? lc[q] : 0) + (int64_t)ks * s2p;
  const int64_t step = (int64_t)nks * s2p;
  const int n0 = (min(u2, JT) - ks + nks - 1) / nks;   // j's of this warp's slice in the first pass (warp-uniform)
  double m[FQR][R];
#pragma unroll
  for (int i = 0; i < FQR; ++i) {
    if (i < n0) {
#pragma unroll
      for (int q = 0; q < R; ++q) m[i][q] = p[q][i * step];
    } else {
#pragma unroll
      for (int q = 0; q < R; ++q) m[i][q] = 0.0;
    }
  }
  if (PDL) griddep_wait();
  if (CG) {
    if (tid == 0) {
      wait_count(dp.fdone + it.f, it.tgt_f * dp.epoch, dp.status);             // z of this front is complete
      if (it.parent >= 0) { wait_count(dp.fdone + it.parent, it.tgt_pf * dp.epoch, dp.status); wait_count(dp.bdone + it.parent, it.tgt_pb * dp.epoch, dp.status); }
    }
    __syncthreads();
  }
  double acc[R][NR];
#pragma unroll
  for (int q = 0; q < R; ++q)
#pragma unroll
    for (int r = 0; r < NR; ++r) acc[q][r] = 0.0;
  for (int j0 = 0; j0 < u2; j0 += JT) {
    const int jn = min(JT, u2 - j0);
    if (j0 > 0) {                             // fronts with more than JT update unknowns: further passes
      __syncthreads();
      o0 = o1 = -1;
      if (tid < jn) o0 = 2 * (int64_t)st[(j0 + tid) >> 1] + (tid & 1);
      if (tid + 256 < jn) o1 = 2 * (int64_t)st[(j0 + tid + 256) >> 1] + (tid & 1);
    }
    if (o0 >= 0) {
      double v[NR];
      ldv<CG, NR>(x, o0, v);
#pragma unroll
      for (int r = 0; r < NR; ++r) sm.xs[tid][r] = v[r];
    }
    if (o1 >= 0) {
      double v[NR];
      ldv<CG, NR>(x, o1, v);
#pragma unroll
      for (int r = 0; r < NR; ++r) sm.xs[tid + 256][r] = v[r];
    }
    __syncthreads();
    const int nk = (jn - ks + nks - 1) / nks;
    int i = 0;
    if (j0 == 0) {
#pragma unroll
      for (int ii = 0; ii < FQR; ++ii) {
        if (ii < nk) {
          double xv[NR];
          lds_row<NR>(sm.xs, ks + ii * nks, xv);
#pragma unroll
          for (int q = 0; q < R; ++q)
#pragma unroll
            for (int r = 0; r < NR; ++r) acc[q][r] = fma(m[ii][q], xv[r], acc[q][r]);
        }
      }
      i = FQR;
    }
    const double* const* pp = p;
#pragma unroll 4
    for (; i < nk; ++i) {
      double mv[R], xv[NR];
#pragma unroll
      for (int q = 0; q < R; ++q) mv[q] = pp[q][(int64_t)j0 * s2p + i * step];
      lds_row<NR>(sm.xs, ks + i * nks, xv);
#pragma unroll
      for (int q = 0; q < R; ++q)
#pragma unroll
        for (int r = 0; r < NR; ++r) acc[q][r] = fma(mv[q], xv[r], acc[q][r]);
    }
  }
#pragma unroll
  for (int q = 0; q < R; ++q)
#pragma unroll
    for (int r = 0; r < NR; ++r) sm.part[warp][q][r][lane] = acc[q][r];
  __syncthreads();
  // final stage: warp w sums the partials of right-hand side w % NR for row q = w / NR (all 8 warps share the work)
  for (int job = warp; job < R * NR; job += 8) {
    const int q = job / NR, r = job % NR;
    if (q * 32 + lane < it.ncols) {
      double t = 0.0;
#pragma unroll
      for (int c = 0; c < nks; ++c) t += sm.part[c][q][r][lane];
      double* xp = x + (it.g0 + it.col0 + q * 32 + lane) * NR + r;
      *xp = ldx<CG>(xp) - t;
    }
  }
}

// backward, fronts with MANY update unknowns: one warp per pivot column of W^T in the packed left block column
// (lanes over the update unknowns), 8 columns per CTA.  Parallel over all 2s x 2u entries, every factor entry in
// flight before the dependency wait — the shape the latency-bound upper levels of the tree need.
constexpr int BWD_COLS = 8;
constexpr int BQ = 8;   // factor entries per lane preloaded before the wait
template <bool CG, int NR, bool PDL>
__device__ __forceinline__ void backward_item_cols(const BwdItem& it, const PlanView& P, double* x, const Deps& dp) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int u2 = it.u2;
  const bool active = warp < it.ncols;
  const int32_t* st = P.strct + it.soff;
  const double* wc = P.pool + it.foff + it.s2 + (int64_t)(it.col0 + warp) * it.ld;   // W^T(j, col) at col*ld + s2 + j
  // ---- static prefetch: factor column and gather offsets
  double wreg[BQ]; int64_t xo[BQ];
#pragma unroll
  for (int q = 0; q < BQ; ++q) {
    const int j = lane + 32 * q;
    const bool ok = active && j < u2;
    wreg[q] = ok ? wc[j] : 0.0;
    xo[q] = ok ? 2 * (int64_t)st[j >> 1] + (j & 1) : -1;
  }
  if (PDL) griddep_wait();
  if (CG) {
    if (threadIdx.x == 0) {
      wait_count(dp.fdone + it.f, it.tgt_f * dp.epoch, dp.status);             // z of this front is complete
      if (it.parent >= 0) { wait_count(dp.fdone + it.parent, it.tgt_pf * dp.epoch, dp.status); wait_count(dp.bdone + it.parent, it.tgt_pb * dp.epoch, dp.status); }
    }
    __syncthreads();
  }
  if (!active) return;
  double a[NR];
#pragma unroll
  for (int r = 0; r < NR; ++r) a[r] = 0.0;
#pragma unroll
  for (int q = 0; q < BQ; ++q) {
    if (xo[q] >= 0) {
      double xv[NR];
      ldv<CG, NR>(x, xo[q], xv);
#pragma unroll
      for (int r = 0; r < NR; ++r) a[r] = fma(wreg[q], xv[r], a[r]);
    }
  }
  for (int j = lane + 32 * BQ; j < u2; j += 32) {
    const double wv = wc[j];
    const int64_t o = 2 * (int64_t)st[j >> 1] + (j & 1);
    double xv[NR];
    ldv<CG, NR>(x, o, xv);
#pragma unroll
    for (int r = 0; r < NR; ++r) a[r] = fma(wv, xv[r], a[r]);
  }
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    double v = a[r];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
    if (lane == 0) { double* xp = x + (it.g0 + it.col0 + warp) * NR + r; *xp = ldx<CG>(xp) - v; }
  }
}

template <bool CG, int NR, bool PDL = false>
__device__ __forceinline__ void backward_item(const BwdItem& it, const PlanView& P, double* x, BwdSmem<NR>& sm,
                                              const Deps& dp) {
  if (it.G == 0) backward_item_cols<CG, NR, PDL>(it, P, x, dp);
  else if (it.G == 2) backward_item_r<CG, NR, 2, PDL>(it, P, x, sm, dp);
  else backward_item_r<CG, NR, 1, PDL>(it, P, x, sm, dp);
}

// ---- warp-per-task sweeps for SMALL fronts -------------------------------------------------------------------
// The CTA-cooperative items above minimise the latency of one large front (k-split, everything prefetched) and are
// what the top of the tree needs.  The bottom of the tree — most fronts of a design, and thousands per launch in a
// forest of designs — are small: a CTA of 256 mostly idle threads per front wastes the SM's thread slots and the
// launch ends up bound by CTA turnover, not by memory.  A small front is served by one WARP per task instead, which
// never synchronises with another warp: a forward task is a chunk of 32*R rows of the packed left block column
// [F11^-1 ; W^T] (lanes over rows), a backward task 32*R pivot columns against the row-major copy of W (lanes over
// pivot columns); both are coalesced matrix-vector products whose contraction index is staged through shared
// memory in tiles, with 16 factor entries per lane in flight.  Which path a front takes depends only on its own
// size, so a design gives bit-identical results alone and inside a forest.
constexpr int KT = 64;     // contraction-index tile staged in shared memory (per warp)

// acc += M[:, 0..kn) * vs[0..kn): 8 factor entries per lane in flight; `pre` holds the first batch when PRE
template <int NR, int R, bool PRE>
__device__ __forceinline__ void warp_gemv_tile(const double* __restrict__ Mk, int64_t ldm, int kn, const bool (&valid)[R],
                                               const double (*vs)[NR], double (&acc)[R][NR], const double (&pre)[8 / R][R]) {
  constexpr int U = 8 / R;
  for (int kk0 = 0; kk0 < kn; kk0 += U) {
    double m[U][R];
    if (PRE && kk0 == 0) {
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int q = 0; q < R; ++q) m[u][q] = pre[u][q];
    } else {
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int q = 0; q < R; ++q) m[u][q] = (kk0 + u < kn && valid[q]) ? Mk[(int64_t)(kk0 + u) * ldm + 32 * q] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (kk0 + u < kn) {
#pragma unroll
        for (int r = 0; r < NR; ++r) {
          const double y = vs[kk0 + u][r];
#pragma unroll
          for (int q = 0; q < R; ++q) acc[q][r] = fma(m[u][q], y, acc[q][r]);
        }
      }
    }
  }
}

template <int NR, int R, bool PDL>
__device__ __forceinline__ void forward_task(const FwdTask& t, const int32_t* __restrict__ gsrc, const double* __restrict__ fac,
                                             const RhsView& rv, double (*ys)[NR], int lane) {
  constexpr int U = 8 / R;
  const int s2 = t.s2, nf2 = t.rows;
  const int64_t ldp = t.ldp;
  const int32_t* g = gsrc + t.goff;
  const double* L = fac + t.lo + t.r0 + lane;
  bool valid[R];
#pragma unroll
  for (int q = 0; q < R; ++q) valid[q] = lane + 32 * q < t.nr;
  // static prefetch: the first batch of factor entries is on its way before the inputs are touched
  double pre[U][R];
#pragma unroll
  for (int u = 0; u < U; ++u)
#pragma unroll
    for (int q = 0; q < R; ++q) pre[u][q] = (u < s2 && valid[q]) ? L[(int64_t)u * ldp + 32 * q] : 0.0;
  if (PDL) griddep_wait();                 // the previous level is complete and visible from here on
  double acc[R][NR];
#pragma unroll
  for (int q = 0; q < R; ++q)
#pragma unroll
    for (int r = 0; r < NR; ++r) acc[q][r] = 0.0;
  for (int k0 = 0; k0 < s2; k0 += KT) {
    const int kn = min(KT, s2 - k0);
    // right-hand side + children's updates of the pivot unknowns k0.., fixed order (rhs + child 0) + child 1 ...
    for (int h = lane; h < kn; h += 32) {
      const int k = k0 + h;
      double v[NR], w[NR];
      ldv<false, NR>(rv.rhs, t.g0 + k, v);
      for (int c = 0; c < t.nch; ++c) {
        const int i = g[c * nf2 + k];
        if (i >= 0) {
          ldv<false, NR>(rv.upd, i, w);
#pragma unroll
          for (int r = 0; r < NR; ++r) v[r] += w[r];
        }
      }
#pragma unroll
      for (int r = 0; r < NR; ++r) ys[h][r] = v[r];
    }
    __syncwarp();
    if (k0 == 0) warp_gemv_tile<NR, R, true>(L, ldp, kn, valid, ys, acc, pre);
    else warp_gemv_tile<NR, R, false>(L + (int64_t)k0 * ldp, ldp, kn, valid, ys, acc, pre);
    __syncwarp();
  }
#pragma unroll
  for (int q = 0; q < R; ++q) {
    if (!valid[q]) continue;
    const int row = t.r0 + lane + 32 * q;
    if (row < s2) {
      stv<NR>(rv.out, t.g0 + row, acc[q]);
    } else {
      double yt[NR], w[NR];
#pragma unroll
      for (int r = 0; r < NR; ++r) yt[r] = 0.0;
      for (int c = 0; c < t.nch; ++c) {
        const int i = g[c * nf2 + row];
        if (i >= 0) {
          ldv<false, NR>(rv.upd, i, w);
#pragma unroll
          for (int r = 0; r < NR; ++r) yt[r] += w[r];
        }
      }
#pragma unroll
      for (int r = 0; r < NR; ++r) yt[r] -= acc[q][r];
      stv<NR>(rv.upd, (int64_t)t.uoff + (row - s2), yt);
    }
  }
}

template <int NR, int R, bool PDL>
__device__ __forceinline__ void backward_task(const BwdTask& t, const int32_t* __restrict__ strct, const double* __restrict__ fac,
                                              double* x, double (*xs)[NR], int lane) {
  constexpr int U = 8 / R;
  const int u2 = t.u2;
  const int64_t s2p = t.s2p;
  const int32_t* st = strct + t.soff;
  const double* W = fac + t.wo + t.c0 + lane;
  bool valid[R];
#pragma unroll
  for (int q = 0; q < R; ++q) valid[q] = lane + 32 * q < t.nc;
  double pre[U][R];
#pragma unroll
  for (int u = 0; u < U; ++u)
#pragma unroll
    for (int q = 0; q < R; ++q) pre[u][q] = (u < u2 && valid[q]) ? W[(int64_t)u * s2p + 32 * q] : 0.0;
  // gather offsets of the first tile are static too
  int64_t o0 = -1, o1 = -1;
  if (lane < u2) o0 = 2 * (int64_t)st[lane >> 1] + (lane & 1);
  if (lane + 32 < u2) o1 = 2 * (int64_t)st[(lane + 32) >> 1] + (lane & 1);
  if (PDL) griddep_wait();
  double acc[R][NR];
#pragma unroll
  for (int q = 0; q < R; ++q)
#pragma unroll
    for (int r = 0; r < NR; ++r) acc[q][r] = 0.0;
  for (int j0 = 0; j0 < u2; j0 += KT) {
    const int jn = min(KT, u2 - j0);
    if (j0 == 0) {
      if (o0 >= 0) {
        double v[NR];
        ldv<false, NR>(x, o0, v);
#pragma unroll
        for (int r = 0; r < NR; ++r) xs[lane][r] = v[r];
      }
      if (o1 >= 0) {
        double v[NR];
        ldv<false, NR>(x, o1, v);
#pragma unroll
        for (int r = 0; r < NR; ++r) xs[lane + 32][r] = v[r];
      }
    } else {
      for (int h = lane; h < jn; h += 32) {
        const int j = j0 + h;
        const int64_t o = 2 * (int64_t)st[j >> 1] + (j & 1);
        double v[NR];
        ldv<false, NR>(x, o, v);
#pragma unroll
        for (int r = 0; r < NR; ++r) xs[h][r] = v[r];
      }
    }
    __syncwarp();
    if (j0 == 0) warp_gemv_tile<NR, R, true>(W, s2p, jn, valid, xs, acc, pre);
    else warp_gemv_tile<NR, R, false>(W + (int64_t)j0 * s2p, s2p, jn, valid, xs, acc, pre);
    __syncwarp();
  }
#pragma unroll
  for (int q = 0; q < R; ++q) {
    if (!valid[q]) continue;
    double v[NR];
    ldv<false, NR>(x, t.g0 + t.c0 + lane + 32 * q, v);
#pragma unroll
    for (int r = 0; r < NR; ++r) v[r] -= acc[q][r];
    stv<NR>(x, t.g0 + t.c0 + lane + 32 * q, v);
  }
}

// large fronts: one CTA-cooperative item per CTA
template <int NR, bool PDL>
__global__ void __launch_bounds__(256, 4) forward_kernel(const FwdItem* __restrict__ items, const int32_t* __restrict__ gsrc, PlanView P,
                                                       RhsView rv) {
  __shared__ SweepSmem<NR> sm;
  if (PDL) griddep_launch_dependents();    // let the next launch start its static prefetch
  const Deps none{nullptr, nullptr, nullptr, 0, nullptr};
  forward_item<false, NR, PDL>(items[blockIdx.x], gsrc, P, rv, sm, none);
}

template <int NR, bool PDL>
__global__ void __launch_bounds__(256, 4) backward_kernel(const BwdItem* __restrict__ items, PlanView P, double* x) {
  __shared__ BwdSmem<NR> sm;
  if (PDL) griddep_launch_dependents();
  const Deps none{nullptr, nullptr, nullptr, 0, nullptr};
  backward_item<false, NR, PDL>(items[blockIdx.x], P, x, sm, none);
}

// Bottom of the elimination forest: ONE CTA per subtree of small fronts (the fronts below level `subtree_levels`).  The
// 8 warps run the warp tasks of the subtree level by level — leaves first on the way up, the subtree's root first on the
// way down — with a CTA barrier between levels: the update vectors / unknowns a level produces are consumed by the same
// CTA (global memory, visible after the barrier), so levels 0..H-1 of a sweep cost one launch instead of H, and the
// dependent round trips of a small front overlap with those of its siblings and cousins in the same CTA.
// sub[b] = {first entry of the subtree's level ranges in ptr, number of levels}; tasks of level l: ptr[o+l] .. ptr[o+l+1].
template <int NR, bool PDL>
__global__ void __launch_bounds__(256, 4) forward_subtree_kernel(const int2* __restrict__ sub, const int32_t* __restrict__ ptr,
                                                                  const FwdTask* __restrict__ tasks, const int32_t* __restrict__ gsrc,
                                                                  const double* __restrict__ fac, RhsView rv) {
  __shared__ __align__(16) double tile[8][KT][NR];
  if (PDL) griddep_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int2 sb = sub[blockIdx.x];
  for (int l = 0; l < sb.y; ++l) {
    const int t1 = ptr[sb.x + l + 1];
    for (int id = ptr[sb.x + l] + warp; id < t1; id += 8) {
      const FwdTask t = tasks[id];
      if (t.nr <= 32) forward_task<NR, 1, PDL>(t, gsrc, fac, rv, tile[warp], lane);
      else forward_task<NR, 2, PDL>(t, gsrc, fac, rv, tile[warp], lane);
    }
    if (l + 1 < sb.y) __syncthreads();
  }
}

template <int NR, bool PDL>
__global__ void __launch_bounds__(256, 4) backward_subtree_kernel(const int2* __restrict__ sub, const int32_t* __restrict__ ptr,
                                                                   const BwdTask* __restrict__ tasks, const int32_t* __restrict__ strct,
                                                                   const double* __restrict__ fac, double* x) {
  __shared__ __align__(16) double tile[8][KT][NR];
  if (PDL) griddep_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int2 sb = sub[blockIdx.x];
  for (int l = sb.y - 1; l >= 0; --l) {
    const int t1 = ptr[sb.x + l + 1];
    for (int id = ptr[sb.x + l] + warp; id < t1; id += 8) {
      const BwdTask t = tasks[id];
      if (t.nc <= 32) backward_task<NR, 1, PDL>(t, strct, fac, x, tile[warp], lane);
      else backward_task<NR, 2, PDL>(t, strct, fac, x, tile[warp], lane);
    }
    if (l > 0) __syncthreads();
  }
}

// ---- pack: the solve phase reads only the left block column of a front (and W a second time, row-major) ----
// Copy both into dense, 32-byte aligned panels; the front pool is factorisation workspace only.
__global__ void __launch_bounds__(256) pack_kernel(PlanView P, const int64_t* __restrict__ lo, const int64_t* __restrict__ wo,
                                                   const int32_t* __restrict__ ldp, double* __restrict__ fac) {
  __shared__ double tile[32][33];
  const int f = blockIdx.x;
  const int s2 = 2 * P.s[f], u2 = 2 * front_u(P, f);
  const int64_t ld = s2 + u2, lp = ldp[f];
  const double* src = P.pool + P.foff[f];
  double* dst = fac + lo[f];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int k = warp; k < s2; k += 8)
    for (int i = lane; i < ld; i += 32) dst[k * lp + i] = src[k * ld + i];
  if (u2 == 0) return;
  double* dw = fac + wo[f];
  const int64_t s2p = (s2 + 3) & ~3;
  for (int c0 = 0; c0 < s2; c0 += 32)
    for (int j0 = 0; j0 < u2; j0 += 32) {
      __syncthreads();
      for (int w = warp; w < 32; w += 8) {              // column c0+w of W^T, rows j0..j0+31
        const int c = c0 + w, j = j0 + lane;
        tile[w][lane] = (c < s2 && j < u2) ? src[(int64_t)c * ld + s2 + j] : 0.0;
      }
      __syncthreads();
      for (int w = warp; w < 32; w += 8) {              // row j0+w of the copy, columns c0..c0+31
        const int j = j0 + w, c = c0 + lane;
        if (j < u2 && c < s2) dw[(int64_t)j * s2p + c] = tile[lane][w];
      }
    }
}

// ---- persistent operator kernel: x = refine((A - sigma B)^-1 b) in ONE cooperative launch ----------------
// As separate launches (even inside a CUDA graph) every step of a sweep pays the launch/drain gap, and with
// several designs in flight the GPU front end becomes the limiter (~3 us per kernel node).  Here a
// co-resident grid runs the whole operator application: items sit in one queue ordered children-before-
// parents; CTA b takes items b, b+G, b+2G... in order and, instead of a grid-wide barrier per level, waits
// only for the fronts it depends on (per-front completion counters in global memory, dataflow).  Items a CTA
// waits for always precede it in the queue and every CTA is resident (cooperative launch), so the wait
// cannot deadlock; a spin limit turns a bug into an error flag instead of a hung GPU.  Grid barriers remain
// only around the refinement SpMV.
struct OpArgs {
  PlanView P;
  const FwdItem* fwd_q; const BwdItem* bwd_q;   // forward queue (levels ascending), backward queue (levels descending)
  const int32_t* gsrc;
  int n_fwd, n_bwd;
  int32_t* fdone; int32_t* bdone; int32_t* status; const int32_t* nfs;
  int epoch0;                                 // sweeps completed before this launch
  const double* b; double* x; double* upd;   // b and x: length 2n, permuted interleaved layout
  double* rt; double* rdx;                    // refinement work vectors
  int refine;
  int32_t n; const int32_t* rowptr; const int32_t* col; const double* vals; int64_t nnz; const double* sigma_node;
};

union OpSmem {
  SweepSmem<1> f;
  BwdSmem<1> b;
};

__device__ __forceinline__ void sweep_dataflow(const OpArgs& a, const double* rhs, double* out, int epoch, OpSmem& sm) {
  const Deps dp{a.fdone, a.bdone, a.status, epoch, a.nfs};
  const RhsView rv{rhs, out, a.upd};
  for (int i = blockIdx.x; i < a.n_fwd; i += gridDim.x) {
    const FwdItem it = a.fwd_q[i];
    forward_item<true, 1>(it, a.gsrc, a.P, rv, sm.f, dp);
    __threadfence();
    __syncthreads();                       // all writes of the item are fenced; shared memory is free again
    if (threadIdx.x == 0) atomicAdd(a.fdone + it.f, 1);
  }
  for (int i = blockIdx.x; i < a.n_bwd; i += gridDim.x) {
    const BwdItem it = a.bwd_q[i];
    backward_item<true, 1>(it, a.P, out, sm.b, dp);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) atomicAdd(a.bdone + it.f, 1);
  }
}

__global__ void __launch_bounds__(256) op_kernel(OpArgs a) {
  __shared__ OpSmem sm;
  cooperative_groups::grid_group grid = cooperative_groups::this_grid();
  sweep_dataflow(a, a.b, a.x, a.epoch0 + 1, sm);
  for (int r = 0; r < a.refine; ++r) {
    grid.sync();                            // x complete everywhere
    // rt = b - K x, four threads per row
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    const int l32 = threadIdx.x & 31;       // whole warps stay in the loop together: the shuffles below use the full mask
    for (int64_t gid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; gid - l32 < 4 * (int64_t)a.n; gid += nthreads) {
      const int64_t row = gid >> 2;
      const int lane = (int)(gid & 3);
      double ax = 0.0, ay = 0.0;
      if (row < a.n) {
        const double sig = a.sigma_node[row];
        for (int32_t z = a.rowptr[row] + lane; z < a.rowptr[row + 1]; z += 4) {
          const double smv = sig * a.vals[(int64_t)S_MINV * a.nnz + z];
          const double* xp = a.x + 2 * (int64_t)a.col[z];
          const double vx = __ldcg(xp), vy = __ldcg(xp + 1);
          ax = fma(a.vals[(int64_t)S_AXX * a.nnz + z] - smv, vx, ax);
          ax = fma(a.vals[(int64_t)S_AXY * a.nnz + z], vy, ax);
          ay = fma(a.vals[(int64_t)S_AYX * a.nnz + z], vx, ay);
          ay = fma(a.vals[(int64_t)S_AYY * a.nnz + z] - smv, vy, ay);
        }
      }
      ax += __shfl_down_sync(0xffffffffu, ax, 2, 4); ay += __shfl_down_sync(0xffffffffu, ay, 2, 4);
      ax += __shfl_down_sync(0xffffffffu, ax, 1, 4); ay += __shfl_down_sync(0xffffffffu, ay, 1, 4);
      if (row < a.n && lane == 0) {
        a.rt[2 * row] = a.b[2 * row] - ax;
        a.rt[2 * row + 1] = a.b[2 * row + 1] - ay;
      }
    }
    grid.sync();                            // rt complete
    sweep_dataflow(a, a.rt, a.rdx, a.epoch0 + 2 + r, sm);
    grid.sync();                            // rdx complete
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < 2 * (int64_t)a.n; i += nthreads) a.x[i] = __ldcg(a.x + i) + __ldcg(a.rdx + i);
  }
}

PlanView view(const DevPlan& D);
PlanView sweep_view(const DevPlan& D) {   // the sweeps read the packed panels, never the front pool
  PlanView v = view(D);
  v.pool = D.fac.p;
  return v;
}

size_t invert_smem(int m) { return ((size_t)(m | 1) * m + 2 * (size_t)m) * sizeof(double); }

PlanView view(const DevPlan& D) {
  PlanView v;
  v.first = D.first.p; v.s = D.s.p; v.sptr = D.sptr.p; v.strct = D.strct.p; v.cptr = D.cptr.p; v.child = D.child.p;
  v.cmap_ptr = D.cmap_ptr.p; v.cmap = D.cmap.p; v.uoff = D.uoff.p; v.foff = D.foff.p; v.pool = D.pool.p;
  return v;
}


}  // namespace

// fronts with at most this many pivot / update unknowns take the warp-per-task path of the sweeps (0: none does)
int small_front_limit(const char* env, int dflt) {
  const char* e = std::getenv(env);
  return e ? std::max(0, atoi(e)) : dflt;
}

void build_dev_plan(plfem_ctx* ctx, const FrontPlan& P, DevPlan& D) {
  if (2 * P.max_s > MAX_PIV) throw StatusError(PLFEM_ERR_INVALID, "max_sn_nodes must be <= 64");
  D.n = P.n; D.nfronts = P.nfronts; D.nlevels = P.nlevels;
  D.first.upload(ctx, P.first); D.s.upload(ctx, P.s); D.sptr.upload(ctx, P.sptr); D.strct.upload(ctx, P.strct);
  D.sn_of.upload(ctx, P.sn_of); D.parent.upload(ctx, P.parent); D.cptr.upload(ctx, P.cptr); D.child.upload(ctx, P.child);
  D.cmap_ptr.upload(ctx, P.cmap_ptr); D.cmap.upload(ctx, P.cmap);
  {
    // fronts of a level in two size classes (pivot block <= 64 unknowns first): the pivot-block inverse runs 256-thread
    // CTAs on the first class and 1024-thread CTAs on the second — one oversized leaf must not put a thousand threads on
    // each of the thousands of 48 x 48 blocks of its level
    std::vector<int32_t> lf(P.lfront);
    D.lsplit.assign(P.nlevels, 0); D.lmax_small.assign(P.nlevels, 0);
    for (int l = 0; l < P.nlevels; ++l) {
      auto mid = std::stable_partition(lf.begin() + P.lptr[l], lf.begin() + P.lptr[l + 1], [&](int32_t f) { return 2 * P.s[f] <= 64; });
      D.lsplit[l] = (int32_t)(mid - (lf.begin() + P.lptr[l]));
      for (auto it = lf.begin() + P.lptr[l]; it != mid; ++it) D.lmax_small[l] = std::max(D.lmax_small[l], 2 * P.s[*it]);
    }
    D.lfront.upload(ctx, lf);
  }
  D.foff.upload(ctx, P.foff);
  D.lptr = P.lptr;
  std::vector<int32_t> uoff(P.nfronts + 1, 0);
  for (int f = 0; f < P.nfronts; ++f) uoff[f + 1] = uoff[f] + 2 * (P.sptr[f + 1] - P.sptr[f]);
  D.upd_len = uoff[P.nfronts];
  D.uoff.upload(ctx, uoff);
  D.upd.alloc(ctx, (size_t)SOLVE_NRHS * std::max<int64_t>(D.upd_len, 1));

  // packed panels of the solve phase
  std::vector<int64_t> lo(P.nfronts + 1, 0), wo(P.nfronts, 0);
  std::vector<int32_t> ldp(P.nfronts, 0);
  {
    int64_t off = 0;
    for (int f = 0; f < P.nfronts; ++f) {
      const int64_t s2 = 2 * (int64_t)P.s[f], u2 = 2 * (int64_t)(P.sptr[f + 1] - P.sptr[f]);
      ldp[f] = (int32_t)((s2 + u2 + 3) & ~int64_t(3));
      lo[f] = off; off += (int64_t)ldp[f] * s2;
      wo[f] = off; off += ((s2 + 3) & ~int64_t(3)) * u2;
    }
    lo[P.nfronts] = off;
    D.fac.alloc(ctx, (size_t)std::max<int64_t>(off, 1));
    D.lo.upload(ctx, lo); D.wo.upload(ctx, wo); D.ldp.upload(ctx, ldp);
  }
  std::vector<int4> wt, stl, ea;
  std::vector<FwdItem> fw; std::vector<BwdItem> bw;
  std::vector<FwdItem> fwb; std::vector<BwdItem> bwb;     // per-level launches: CTA items of the fronts above the subtrees
  D.fwdb_ptr.assign(P.nlevels + 1, 0); D.bwdb_ptr = D.fwdb_ptr;
  // The fronts below level H (PLFEM_SUBTREE_LEVELS, default 1: the leaves — hundreds per design, ~75 x 30 entries each; a
  // 256-thread CTA per leaf is bound by CTA turnover) are served by warp tasks, one CTA per maximal subtree of such fronts.
  const int H = small_front_limit("PLFEM_SUBTREE_LEVELS", 1);
  auto in_sub = [&](int f) { return P.level[f] < H; };
  D.w_ptr.assign(P.nlevels + 1, 0); D.s_ptr = D.ea_ptr = D.fwd_ptr = D.bwd_ptr = D.w_ptr;
  D.lmax_m.assign(P.nlevels, 0);
  // flattened child -> parent gather: per front, for every front row (2nf unknowns) the offset into the
  // update-vector pool it receives from the first and from the second child (-1 = nothing)
  std::vector<int32_t> goff(P.nfronts + 1, 0);
  for (int f = 0; f < P.nfronts; ++f)   // one table per child, at least two (the CTA path reads two unconditionally)
    goff[f + 1] = goff[f] + 2 * std::max(2, P.cptr[f + 1] - P.cptr[f]) * (P.s[f] + P.sptr[f + 1] - P.sptr[f]);
  std::vector<int32_t> gsrc(goff[P.nfronts], -1);
  for (int f = 0; f < P.nfronts; ++f) {
    const int nf2 = 2 * (P.s[f] + P.sptr[f + 1] - P.sptr[f]);
    for (int q = P.cptr[f]; q < P.cptr[f + 1]; ++q) {
      const int ch = P.child[q];
      int32_t* g = gsrc.data() + goff[f] + (q - P.cptr[f]) * nf2;
      for (int k = P.sptr[ch], o = P.cmap_ptr[ch], idx = 0; k < P.sptr[ch + 1]; ++k, ++o, ++idx) {
        g[2 * P.cmap[o]] = uoff[ch] + 2 * idx;
        g[2 * P.cmap[o] + 1] = uoff[ch] + 2 * idx + 1;
      }
    }
  }
  const int bwd_rows_u2 = small_front_limit("PLFEM_BWD_ROWS_U2", 128);
  std::vector<int32_t> nfs(P.nfronts, 0), nbs(P.nfronts, 0);
  for (int f = 0; f < P.nfronts; ++f) {
    const int s2 = 2 * P.s[f], u2 = 2 * (P.sptr[f + 1] - P.sptr[f]);
    const int rows = s2 + u2;
    nfs[f] = rows <= 128 ? 1 : (rows + 63) / 64;     // forward CTA items: one slab of <= 128 rows, else slabs of 64
    // backward CTA items: few update unknowns -> chunks of 64 pivot columns sharing one gather, many -> 8 columns each
    nbs[f] = u2 == 0 ? 0 : (u2 <= bwd_rows_u2 ? (s2 + 63) / 64 : (s2 + BWD_COLS - 1) / BWD_COLS);
  }
  for (int l = 0; l < P.nlevels; ++l) {
    for (int q = P.lptr[l]; q < P.lptr[l + 1]; ++q) {
      const int f = P.lfront[q];
      const int s = P.s[f], u = P.sptr[f + 1] - P.sptr[f];
      const int s2 = 2 * s, u2 = 2 * u, nf = s + u;
      D.lmax_m[l] = std::max(D.lmax_m[l], s2);
      for (int j0 = 0; j0 < u2; j0 += GT)
        for (int i0 = 0; i0 < s2; i0 += GT) wt.push_back(make_int4(f, j0, i0, 0));
      for (int b0 = 0; b0 < u2; b0 += GT)
        for (int a0 = 0; a0 < u2; a0 += GT) stl.push_back(make_int4(f, a0, b0, 0));
      if (P.cptr[f + 1] > P.cptr[f])
        for (int c0 = 0; c0 < nf; c0 += EA_COLS) ea.push_back(make_int4(f, c0, std::min(c0 + EA_COLS, nf), 0));
      {
        const int rows = s2 + u2;
        // row groups x rows per thread: <= 32 rows (1,1), <= 64 (1,2), <= 128 (2,2), larger fronts in slabs of 64 rows (1,2)
        const int G = (rows > 64 && rows <= 128) ? 2 : 1, Rr = rows <= 32 ? 1 : 2;
        const int nch = P.cptr[f + 1] - P.cptr[f];
        const bool fsmall = in_sub(f);
        for (int r0 = 0; r0 < rows; r0 += 32 * G * Rr) {
          FwdItem it{};
          it.f = f; it.row0 = r0; it.nrows = std::min(32 * G * Rr, rows - r0); it.G = G | (Rr << 8); it.s2 = s2; it.ld = ldp[f];
          it.ch0 = nch > 0 ? P.child[P.cptr[f]] : -1; it.ch1 = nch > 1 ? P.child[P.cptr[f] + 1] : -1;
          it.tgt0 = it.ch0 >= 0 ? nfs[it.ch0] : 0; it.tgt1 = it.ch1 >= 0 ? nfs[it.ch1] : 0;
          it.uoff = uoff[f]; it.goff = goff[f]; it.nchild = nch; it.nf2 = rows;
          it.foff = lo[f]; it.g0 = 2 * (int64_t)P.first[f];
          fw.push_back(it);
          if (!fsmall) fwb.push_back(it);
        }
      }
      if (u2 > 0) {
        const bool bsmall = in_sub(f);
        const bool rows_style = u2 <= bwd_rows_u2;
        const int Gb = rows_style ? (s2 <= 32 ? 1 : 2) : 0;   // pivot columns per thread; 0 = one warp per column
        const int cw = rows_style ? 32 * Gb : BWD_COLS;
        for (int c0 = 0; c0 < s2; c0 += cw) {
          BwdItem it{};
          const int pa = P.parent[f];
          it.f = f; it.col0 = c0; it.ncols = std::min(cw, s2 - c0); it.s2 = s2; it.u2 = u2; it.soff = P.sptr[f];
          it.ld = rows_style ? ((s2 + 3) & ~3) : ldp[f];
          it.parent = pa; it.tgt_f = nfs[f]; it.tgt_pf = pa >= 0 ? nfs[pa] : 0; it.tgt_pb = pa >= 0 ? nbs[pa] : 0; it.G = Gb;
          it.foff = rows_style ? wo[f] : lo[f]; it.g0 = 2 * (int64_t)P.first[f];
          bw.push_back(it);
          if (!bsmall) bwb.push_back(it);
        }
      }
    }
    D.w_ptr[l + 1] = (int32_t)wt.size(); D.s_ptr[l + 1] = (int32_t)stl.size(); D.ea_ptr[l + 1] = (int32_t)ea.size();
    D.fwd_ptr[l + 1] = (int32_t)fw.size(); D.bwd_ptr[l + 1] = (int32_t)bw.size();
    D.fwdb_ptr[l + 1] = (int32_t)fwb.size(); D.bwdb_ptr[l + 1] = (int32_t)bwb.size();
  }
  D.fwdb_items.upload(ctx, fwb); D.bwdb_items.upload(ctx, bwb);
  {
    // subtree task lists.  Fronts are numbered in post-order: the subtree of root r is the contiguous range ending at r.
    std::vector<int32_t> size(P.nfronts, 1);
    for (int f = 0; f < P.nfronts; ++f) if (P.parent[f] >= 0) size[P.parent[f]] += size[f];
    std::vector<FwdTask> ft; std::vector<BwdTask> bt;
    std::vector<int32_t> fptr(1, 0), bptr(1, 0);
    std::vector<int2> subs;
    // a CTA serves a GROUP of neighbouring subtrees, level by level, so that its 8 warps have tasks at the widest level
    std::vector<std::vector<int32_t>> by_level;
    auto close_group = [&] {
      if (by_level.empty()) return;
      subs.push_back(make_int2((int)fptr.size() - 1, (int)by_level.size()));
      for (const std::vector<int32_t>& fl : by_level) {
        for (int f : fl) {
          const int s2 = 2 * P.s[f], u2 = 2 * (P.sptr[f + 1] - P.sptr[f]), rows = s2 + u2;
          const int nch = P.cptr[f + 1] - P.cptr[f];
          // rows per warp: at most 64 factor entries per lane in flight order, so a task is a handful of round trips
          const int Rf = (s2 <= 32 && rows > 32) ? 2 : 1;
          for (int r0 = 0; r0 < rows; r0 += 32 * Rf) {
            FwdTask t{};
            t.s2 = s2; t.rows = rows; t.r0 = r0; t.nr = std::min(32 * Rf, rows - r0); t.ldp = ldp[f]; t.goff = goff[f]; t.uoff = uoff[f];
            t.nch = nch; t.lo = lo[f]; t.g0 = 2 * (int64_t)P.first[f];
            ft.push_back(t);
          }
          if (u2 > 0) {
            const int Rb = (s2 > 32 && u2 <= 32) ? 2 : 1;
            for (int c0 = 0; c0 < s2; c0 += 32 * Rb) {
              BwdTask t{};
              t.s2 = s2; t.u2 = u2; t.c0 = c0; t.nc = std::min(32 * Rb, s2 - c0); t.s2p = (s2 + 3) & ~3; t.soff = P.sptr[f];
              t.wo = wo[f]; t.g0 = 2 * (int64_t)P.first[f];
              bt.push_back(t);
            }
          }
        }
        fptr.push_back((int32_t)ft.size()); bptr.push_back((int32_t)bt.size());
      }
      by_level.clear();
    };
    int widest = 0;     // forward tasks of the current group at its widest level
    for (int r = 0; r < P.nfronts; ++r) {
      if (!in_sub(r) || (P.parent[r] >= 0 && in_sub(P.parent[r]))) continue;     // not the root of a maximal subtree
      const int nl = P.level[r] + 1;
      if ((int)by_level.size() < nl) by_level.resize(nl);
      for (int f = r - size[r] + 1; f <= r; ++f) by_level[P.level[f]].push_back(f);
      widest = 0;
      for (const std::vector<int32_t>& fl : by_level) {
        int nt = 0;
        for (int f : fl) { const int s2 = 2 * P.s[f], rows = s2 + 2 * (P.sptr[f + 1] - P.sptr[f]); nt += (rows + ((s2 <= 32 && rows > 32) ? 63 : 31)) / ((s2 <= 32 && rows > 32) ? 64 : 32); }
        widest = std::max(widest, nt);
      }
      if (widest >= 8) close_group();
    }
    close_group();
    D.n_subs = (int)subs.size();
    D.fwd_tasks.upload(ctx, ft); D.bwd_tasks.upload(ctx, bt);
    D.sub_fptr.upload(ctx, fptr); D.sub_bptr.upload(ctx, bptr); D.subs.upload(ctx, subs);
  }
  {
    // backward queue of the persistent operator kernel: levels descending; completion counters
    std::vector<BwdItem> bq; bq.reserve(bw.size());
    for (int l = P.nlevels - 1; l >= 0; --l) bq.insert(bq.end(), bw.begin() + D.bwd_ptr[l], bw.begin() + D.bwd_ptr[l + 1]);
    D.bwd_q.upload(ctx, bq);
    D.n_fwd = (int)fw.size(); D.n_bwd = (int)bq.size();
    D.fdone.alloc(ctx, P.nfronts); D.bdone.alloc(ctx, P.nfronts);
    D.nfs.upload(ctx, nfs);
  }
  D.gsrc.upload(ctx, gsrc);
  D.w_tiles.upload(ctx, wt); D.s_tiles.upload(ctx, stl); D.ea_slabs.upload(ctx, ea);
  D.fwd_items.upload(ctx, fw); D.bwd_items.upload(ctx, bw);
  D.pool.alloc(ctx, (size_t)P.foff[P.nfronts]);
  D.status.alloc(ctx, 4);
  // the host vectors above are pageable: make sure the copies are done before they go out of scope
  PLFEM_CUDA(stream_wait(ctx->stream));
}

void launch_front_load(plfem_ctx* ctx, const DevPattern& pat, DevPlan& D, const double* d_vals, const double* d_sigma_node) {
  PLFEM_CUDA(cudaMemsetAsync(D.pool.p, 0, D.pool.n * sizeof(double), ctx->stream));
  PLFEM_CUDA(cudaMemsetAsync(D.status.p, 0, 4 * sizeof(int32_t), ctx->stream));
  PLFEM_CUDA(cudaMemsetAsync(D.fdone.p, 0, D.fdone.n * sizeof(int32_t), ctx->stream));
  PLFEM_CUDA(cudaMemsetAsync(D.bdone.p, 0, D.bdone.n * sizeof(int32_t), ctx->stream));
  D.epoch = 0;
  const int bs = 256;
  front_load_kernel<<<(unsigned)((pat.nnz + bs - 1) / bs), bs, 0, ctx->stream>>>(pat.nnz, pat.rowidx.p, pat.col.p,
                                                                                 D.sn_of.p, view(D), d_vals, d_sigma_node);
  PLFEM_CUDA(cudaGetLastError());
  ctx->launches++;
}

void run_factorization(plfem_ctx* ctx, const DevPlan& D) {
  static bool attr_set[64] = {};
  if (!(ctx->device < 64 && attr_set[ctx->device])) {
    PLFEM_CUDA(cudaFuncSetAttribute(invert_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)invert_smem(MAX_PIV)));
    if (ctx->device < 64) attr_set[ctx->device] = true;
  }
  const PlanView v = view(D);
  for (int l = 0; l < D.nlevels; ++l) {
    const int nea = D.ea_ptr[l + 1] - D.ea_ptr[l];
    if (nea > 0) {
      extend_add_kernel<<<nea, 256, 0, ctx->stream>>>(D.ea_slabs.p + D.ea_ptr[l], v);
      ctx->launches++;
    }
    const int nfl = D.lptr[l + 1] - D.lptr[l];
    // two size classes in separate launches only where the level is throughput-bound (thousands of small blocks, as at the
    // bottom of a forest); a latency-bound level runs both classes in one launch, concurrently
    const int nsmall = D.lsplit[l];
    if (nsmall >= 1500 && nfl > nsmall) {
      invert_kernel<<<nsmall, 256, invert_smem(D.lmax_small[l]), ctx->stream>>>(D.lfront.p + D.lptr[l], v, D.status.p);
      invert_kernel<<<nfl - nsmall, 1024, invert_smem(D.lmax_m[l]), ctx->stream>>>(D.lfront.p + D.lptr[l] + nsmall, v, D.status.p);
      ctx->launches += 2;
    } else {
      invert_kernel<<<nfl, D.lmax_m[l] > 64 ? 1024 : 256, invert_smem(D.lmax_m[l]), ctx->stream>>>(D.lfront.p + D.lptr[l], v, D.status.p);
      ctx->launches++;
    }
    const int nw = D.w_ptr[l + 1] - D.w_ptr[l];
    if (nw > 0) {
      gemm_w_kernel<<<nw, 256, 0, ctx->stream>>>(D.w_tiles.p + D.w_ptr[l], v);
      ctx->launches++;
    }
    const int ns = D.s_ptr[l + 1] - D.s_ptr[l];
    if (ns > 0) {
      gemm_schur_kernel<<<ns, 256, 0, ctx->stream>>>(D.s_tiles.p + D.s_ptr[l], v);
      ctx->launches++;
    }
  }
  pack_kernel<<<D.nfronts, 256, 0, ctx->stream>>>(v, D.lo.p, D.wo.p, D.ldp.p, D.fac.p);
  ctx->launches++;
  PLFEM_CUDA(cudaGetLastError());
}

bool use_pdl() {
  static const bool on = [] { const char* e = std::getenv("PLFEM_PDL"); return !(e && e[0] == '0'); }();
  return on;
}

template <class... KArgs, class... Args>
void launch_sweep(void (*kernel)(KArgs...), bool pdl, int grid, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 0; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
  PLFEM_CUDA(cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...));
}

// nrhs right-hand sides (1 or SOLVE_NRHS), INTERLEAVED: entry i of right-hand side r at b[i * nrhs + r].  One launch for the
// bottom subtrees, then one launch per remaining level (CTA items).
void run_solve_forward(plfem_ctx* ctx, const DevPlan& D, const double* b, double* z, int nrhs) {
  const PlanView v = sweep_view(D);
  if (nrhs != 1 && nrhs != SOLVE_NRHS) throw StatusError(PLFEM_ERR_INTERNAL, "unsupported number of right-hand sides");
  const RhsView rv{b, z, D.upd.p};
  const bool pdl = use_pdl();
  bool first = true;     // the first launch follows kernels that are not PDL-aware: plain launch
  const int32_t* gs = D.gsrc.p;
  if (D.n_subs > 0) {
    if (nrhs == 1) launch_sweep(forward_subtree_kernel<1, false>, false, D.n_subs, ctx->stream, D.subs.p, D.sub_fptr.p, D.fwd_tasks.p, gs, D.fac.p, rv);
    else launch_sweep(forward_subtree_kernel<SOLVE_NRHS, false>, false, D.n_subs, ctx->stream, D.subs.p, D.sub_fptr.p, D.fwd_tasks.p, gs, D.fac.p, rv);
    first = false; ctx->launches++;
  }
  for (int l = 0; l < D.nlevels; ++l) {
    const int nbig = D.fwdb_ptr[l + 1] - D.fwdb_ptr[l];
    if (nbig == 0) continue;
    const FwdItem* items = D.fwdb_items.p + D.fwdb_ptr[l];
    const bool p = pdl && !first;
    if (nrhs == 1) { if (p) launch_sweep(forward_kernel<1, true>, true, nbig, ctx->stream, items, gs, v, rv); else launch_sweep(forward_kernel<1, false>, false, nbig, ctx->stream, items, gs, v, rv); }
    else { if (p) launch_sweep(forward_kernel<SOLVE_NRHS, true>, true, nbig, ctx->stream, items, gs, v, rv); else launch_sweep(forward_kernel<SOLVE_NRHS, false>, false, nbig, ctx->stream, items, gs, v, rv); }
    first = false; ctx->launches++;
  }
}

// must follow run_solve_forward on the same stream (its launches may be PDL-chained to the forward ones)
void run_solve_backward(plfem_ctx* ctx, const DevPlan& D, double* x, int nrhs) {
  const PlanView v = sweep_view(D);
  const bool pdl = use_pdl();
  for (int l = D.nlevels - 1; l >= 0; --l) {
    const int nbig = D.bwdb_ptr[l + 1] - D.bwdb_ptr[l];
    if (nbig == 0) continue;
    const BwdItem* items = D.bwdb_items.p + D.bwdb_ptr[l];
    if (nrhs == 1) { if (pdl) launch_sweep(backward_kernel<1, true>, true, nbig, ctx->stream, items, v, x); else launch_sweep(backward_kernel<1, false>, false, nbig, ctx->stream, items, v, x); }
    else { if (pdl) launch_sweep(backward_kernel<SOLVE_NRHS, true>, true, nbig, ctx->stream, items, v, x); else launch_sweep(backward_kernel<SOLVE_NRHS, false>, false, nbig, ctx->stream, items, v, x); }
    ctx->launches++;
  }
  if (D.n_subs > 0) {
    if (nrhs == 1) { if (pdl) launch_sweep(backward_subtree_kernel<1, true>, true, D.n_subs, ctx->stream, D.subs.p, D.sub_bptr.p, D.bwd_tasks.p, D.strct.p, D.fac.p, x); else launch_sweep(backward_subtree_kernel<1, false>, false, D.n_subs, ctx->stream, D.subs.p, D.sub_bptr.p, D.bwd_tasks.p, D.strct.p, D.fac.p, x); }
    else { if (pdl) launch_sweep(backward_subtree_kernel<SOLVE_NRHS, true>, true, D.n_subs, ctx->stream, D.subs.p, D.sub_bptr.p, D.bwd_tasks.p, D.strct.p, D.fac.p, x); else launch_sweep(backward_subtree_kernel<SOLVE_NRHS, false>, false, D.n_subs, ctx->stream, D.subs.p, D.sub_bptr.p, D.bwd_tasks.p, D.strct.p, D.fac.p, x); }
    ctx->launches++;
  }
}

int op_grid_size(plfem_ctx* ctx, int ctas_per_sm) {
  static int max_per_sm[64] = {}; static int nsm[64] = {};
  const int dev = ctx->device < 64 ? ctx->device : 0;
  if (!nsm[dev]) {
    int occ = 0, sms = 0;
    PLFEM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, op_kernel, 256, 0));
    PLFEM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device));
    max_per_sm[dev] = std::max(occ, 1); nsm[dev] = sms;
  }
  return nsm[dev] * std::max(1, std::min(ctas_per_sm, max_per_sm[dev]));
}

// x = (A - sigma B)^-1 b with `refine` refinement steps, one cooperative launch (b, x, rt, rdx distinct)
void run_operator(plfem_ctx* ctx, const DevPattern& pat, DevPlan& D, const double* d_vals, const double* d_sigma_node, const double* b,
                  double* x, double* rt, double* rdx, int refine, int ctas_per_sm) {
  OpArgs a;
  a.P = sweep_view(D);
  a.fwd_q = D.fwd_items.p; a.bwd_q = D.bwd_q.p; a.gsrc = D.gsrc.p; a.n_fwd = D.n_fwd; a.n_bwd = D.n_bwd;
  a.fdone = D.fdone.p; a.bdone = D.bdone.p; a.status = D.status.p; a.nfs = D.nfs.p;
  a.epoch0 = D.epoch; D.epoch += 1 + refine;
  a.b = b; a.x = x; a.upd = D.upd.p; a.rt = rt; a.rdx = rdx; a.refine = refine;
  a.n = pat.n; a.rowptr = pat.rowptr.p; a.col = pat.col.p; a.vals = d_vals; a.nnz = pat.nnz; a.sigma_node = d_sigma_node;
  void* args[] = {&a};
  PLFEM_CUDA(cudaLaunchCooperativeKernel((const void*)op_kernel, dim3(op_grid_size(ctx, ctas_per_sm)), dim3(256), args, 0, ctx->stream));
  ctx->launches++;
}

void run_solve(plfem_ctx* ctx, const DevPlan& D, const double* b, double* x, int nrhs) {
  run_solve_forward(ctx, D, b, x, nrhs);
  run_solve_backward(ctx, D, x, nrhs);
}

}  // namespace plfem
