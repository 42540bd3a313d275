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

// ---- pack: the solve phase reads only the left block column of a front (and W a second time, row-major) ----
// Copy both into dense, 32-byte aligned panels; the front pool is factorisation workspace only.
__global__ void __launch_bounds__(256) pack_kernel(PlanView P, const int64_t* __restrict__ lo, const int64_t* __restrict__ wo,
                                                   const int32_t* __restrict__ ldp, const uint8_t* __restrict__ in_sub,
                                                   double* __restrict__ fac) {
  __shared__ double tile[32][33];
  const int f = blockIdx.x;
  if (in_sub[f]) return;          // packed into the streams of the bottom subtrees instead
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

  // bottom subtrees: TMA-streamed, one warp each (sweep_stream.cu); everything above them: one launch per level
  std::vector<uint8_t> in_sub;
  build_stream_plan(ctx, P, uoff, in_sub, D.st);
  D.in_sub.upload(ctx, in_sub);

  // packed panels of the solve phase (fronts above the bottom subtrees)
  std::vector<int64_t> lo(P.nfronts + 1, 0), wo(P.nfronts, 0);
  std::vector<int32_t> ldp(P.nfronts, 0);
  {
    int64_t off = 0;
    for (int f = 0; f < P.nfronts; ++f) {
      const int64_t s2 = 2 * (int64_t)P.s[f], u2 = 2 * (int64_t)(P.sptr[f + 1] - P.sptr[f]);
      ldp[f] = (int32_t)((s2 + u2 + 3) & ~int64_t(3));
      lo[f] = off; if (!in_sub[f]) off += (int64_t)ldp[f] * s2;
      wo[f] = off; if (!in_sub[f]) off += ((s2 + 3) & ~int64_t(3)) * u2;
    }
    lo[P.nfronts] = off;
    D.fac.alloc(ctx, (size_t)std::max<int64_t>(off, 1));
    D.lo.upload(ctx, lo); D.wo.upload(ctx, wo); D.ldp.upload(ctx, ldp);
  }
  std::vector<int4> wt, stl, ea;
  std::vector<FwdItem> fwb; std::vector<BwdItem> bwb;     // per-level launches: CTA items of the fronts above the subtrees
  D.fwdb_ptr.assign(P.nlevels + 1, 0); D.bwdb_ptr = D.fwdb_ptr;
  D.w_ptr.assign(P.nlevels + 1, 0); D.s_ptr = D.ea_ptr = D.w_ptr;
  D.lmax_m.assign(P.nlevels, 0);
  // flattened child -> parent gather: per front above the subtrees, for every front row (2nf unknowns) the offset into the
  // update-vector pool it receives from the first and from the second child (-1 = nothing)
  std::vector<int32_t> goff(P.nfronts + 1, 0);
  for (int f = 0; f < P.nfronts; ++f)   // one table per child, at least two (the CTA path reads two unconditionally)
    goff[f + 1] = goff[f] + (in_sub[f] ? 0 : 2 * std::max(2, P.cptr[f + 1] - P.cptr[f]) * (P.s[f] + P.sptr[f + 1] - P.sptr[f]));
  std::vector<int32_t> gsrc(goff[P.nfronts], -1);
  for (int f = 0; f < P.nfronts; ++f) {
    if (in_sub[f]) continue;
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
  static const int bwd_rows_u2 = [] { const char* e = std::getenv("PLFEM_BWD_ROWS_U2"); return e ? std::max(0, atoi(e)) : 128; }();
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
      if (in_sub[f]) continue;
      {
        const int rows = s2 + u2;
        // row groups x rows per thread: <= 32 rows (1,1), <= 64 (1,2), <= 128 (2,2), larger fronts in slabs of 64 rows (1,2)
        const int G = (rows > 64 && rows <= 128) ? 2 : 1, Rr = rows <= 32 ? 1 : 2;
        const int nch = P.cptr[f + 1] - P.cptr[f];
        for (int r0 = 0; r0 < rows; r0 += 32 * G * Rr) {
          FwdItem it{};
          it.f = f; it.row0 = r0; it.nrows = std::min(32 * G * Rr, rows - r0); it.G = G | (Rr << 8); it.s2 = s2; it.ld = ldp[f];
          it.ch0 = nch > 0 ? P.child[P.cptr[f]] : -1; it.ch1 = nch > 1 ? P.child[P.cptr[f] + 1] : -1;
          it.uoff = uoff[f]; it.goff = goff[f]; it.nchild = nch; it.nf2 = rows;
          it.foff = lo[f]; it.g0 = 2 * (int64_t)P.first[f];
          fwb.push_back(it);
        }
      }
      if (u2 > 0) {
        const bool rows_style = u2 <= bwd_rows_u2;
        const int Gb = rows_style ? (s2 <= 32 ? 1 : 2) : 0;   // pivot columns per thread; 0 = one warp per column
        const int cw = rows_style ? 32 * Gb : BWD_COLS;
        for (int c0 = 0; c0 < s2; c0 += cw) {
          BwdItem it{};
          it.f = f; it.col0 = c0; it.ncols = std::min(cw, s2 - c0); it.s2 = s2; it.u2 = u2; it.soff = P.sptr[f];
          it.ld = rows_style ? ((s2 + 3) & ~3) : ldp[f];
          it.parent = P.parent[f]; it.G = Gb;
          it.foff = rows_style ? wo[f] : lo[f]; it.g0 = 2 * (int64_t)P.first[f];
          bwb.push_back(it);
        }
      }
    }
    D.w_ptr[l + 1] = (int32_t)wt.size(); D.s_ptr[l + 1] = (int32_t)stl.size(); D.ea_ptr[l + 1] = (int32_t)ea.size();
    D.fwdb_ptr[l + 1] = (int32_t)fwb.size(); D.bwdb_ptr[l + 1] = (int32_t)bwb.size();
  }
  D.fwdb_items.upload(ctx, fwb); D.bwdb_items.upload(ctx, bwb);
  D.gsrc.upload(ctx, gsrc);
  D.w_tiles.upload(ctx, wt); D.s_tiles.upload(ctx, stl); D.ea_slabs.upload(ctx, ea);
  D.pool.alloc(ctx, (size_t)P.foff[P.nfronts]);
  D.status.alloc(ctx, 4);
  // the host vectors above are pageable: make sure the copies are done before they go out of scope
  PLFEM_CUDA(stream_wait(ctx->stream));
}

void launch_front_load(plfem_ctx* ctx, const DevPattern& pat, DevPlan& D, const double* d_vals, const double* d_sigma_node) {
  PLFEM_CUDA(cudaMemsetAsync(D.pool.p, 0, D.pool.n * sizeof(double), ctx->stream));
  PLFEM_CUDA(cudaMemsetAsync(D.status.p, 0, 4 * sizeof(int32_t), ctx->stream));
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
  pack_kernel<<<D.nfronts, 256, 0, ctx->stream>>>(v, D.lo.p, D.wo.p, D.ldp.p, D.in_sub.p, D.fac.p);
  ctx->launches++;
  launch_stream_pack(ctx, D);
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
  if (D.st.n_subs > 0) {
    launch_stream_forward(ctx, D, b, z, nrhs, false);
    first = false;
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
  bool any_level = false;     // the stream kernel may be chained by PDL only to a PDL-aware predecessor
  for (int l = D.nlevels - 1; l >= 0; --l) {
    const int nbig = D.bwdb_ptr[l + 1] - D.bwdb_ptr[l];
    if (nbig == 0) continue;
    any_level = true;
    const BwdItem* items = D.bwdb_items.p + D.bwdb_ptr[l];
    if (nrhs == 1) { if (pdl) launch_sweep(backward_kernel<1, true>, true, nbig, ctx->stream, items, v, x); else launch_sweep(backward_kernel<1, false>, false, nbig, ctx->stream, items, v, x); }
    else { if (pdl) launch_sweep(backward_kernel<SOLVE_NRHS, true>, true, nbig, ctx->stream, items, v, x); else launch_sweep(backward_kernel<SOLVE_NRHS, false>, false, nbig, ctx->stream, items, v, x); }
    ctx->launches++;
  }
  if (D.st.n_subs > 0) launch_stream_backward(ctx, D, x, nrhs, pdl && any_level);
}

void run_solve(plfem_ctx* ctx, const DevPlan& D, const double* b, double* x, int nrhs) {
  run_solve_forward(ctx, D, b, x, nrhs);
  run_solve_backward(ctx, D, x, nrhs);
}

}  // namespace plfem
