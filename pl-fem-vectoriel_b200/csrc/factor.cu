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

// ---- pivot block inverse: Gauss-Jordan with partial pivoting, the block in REGISTERS, one CTA per front -------------
// The first version kept the block in shared memory and needed three CTA barriers and three passes over the block per
// elimination step: 430 us of pure latency for a 128 x 128 block, the single largest kernel of a forest once the sweeps had
// been streamed (14 % of the kernel time).  Here thread (i, jg) owns row i and the columns jg, jg + NG, ... (E of them) in
// registers; a step publishes only the pivot row and the pivot column (two barriers, ~3 KB of shared traffic), every warp
// finds the pivot redundantly from the published column, and rows are never swapped: step k pivots on the largest entry of
// column k among the rows not used before (the same choices as partial pivoting with swaps, the same arithmetic), and the
// row / column permutation is undone when the result is written back: A^-1[r][c] = a[p_r][step at which row c was used].
template <int MP, int NG, int E>
__global__ void __launch_bounds__(MP * NG, MP == 128 ? 1 : (MP == 64 ? 3 : 6)) invert_kernel(const int32_t* __restrict__ fronts, PlanView P, int32_t* status) {
  extern __shared__ double sB[];          // the finished block for the symmetrised write-back
  __shared__ __align__(16) double prow[MP];
  __shared__ double cand[MP], pcol[2 * MP], sinv;
  __shared__ int p_of[MP], step_of[MP];
  const int f = fronts[blockIdx.x];
  const int m = 2 * P.s[f];
  const int64_t ld = 2 * (int64_t)(P.s[f] + front_u(P, f));
  double* F = P.pool + P.foff[f];
  const int tid = threadIdx.x, lane = tid & 31;
  const int i = tid % MP, jg = tid / MP;
  // thread (i, jg) owns row i and the E CONTIGUOUS columns jg E ... jg E + E - 1: the column of step k = g E + kt belongs to
  // group g at the STATIC register index kt once the step loop is unrolled over kt (the first version dealt the columns
  // round-robin and paid ~200 select instructions per warp and step to read and patch a[k / NG] without a dynamic index)
  double a[E];
#pragma unroll
  for (int t = 0; t < E; ++t) {
    const int j = jg * E + t;
    a[t] = (i < m && j < m) ? F[(int64_t)j * ld + i] : 0.0;
  }
  bool used = false;
  // A step costs every thread E (load, FMA) pairs and little else: the scaled pivot row is published by the threads that own
  // it, column k is patched afterwards by the one column group that owns it (jg is uniform within a warp), rows and columns
  // beyond m hold zeros and take part without predicates, 1 / pivot is computed by the four threads of the pivot row only.
  for (int g = 0; g < NG; ++g) {
#pragma unroll
    for (int kt = 0; kt < E; ++kt) {
      const int k = g * E + kt;
      if (k >= m) break;                      // uniform
      double* pc = pcol + (k & 1) * MP;       // double-buffered: rows of the previous step may still be reading theirs
      if (jg == g) {
        // an unused row always beats a used one, also when its entry is NaN (-0.5: it loses against every number, the step is
        // flagged as singular, and the pivot order stays a permutation)
        const double mine = a[kt];
        const double am = fabs(mine);
        cand[i] = (i < m && !used) ? (am == am ? am : -0.5) : -1.0;
        pc[i] = mine;
      }
      __syncthreads();
      double best = -1.0; int bi = 0;
#pragma unroll
      for (int r = lane; r < MP; r += 32) {
        const double v = cand[r];
        if (v > best) { best = v; bi = r; }
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, off);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
      }
      const int p = bi;                       // the same in every warp
      if (tid == 0) {
        p_of[k] = p; step_of[p] = k;
        if (!(best > 0.0) || !isfinite(best)) { atomicExch(status, 1); atomicExch(status + 3, f); }
      }
      if (i == p) {
        used = true;
        const double inv = 1.0 / pc[p];
        if (jg == g) sinv = inv;
#pragma unroll
        for (int t = 0; t < E; ++t) prow[jg * E + t] = a[t] * inv;
      }
      __syncthreads();
      const double ci = pc[i];
      const double2* pr = reinterpret_cast<const double2*>(prow + jg * E);
      if (i == p) {
#pragma unroll
        for (int t = 0; t < E; t += 2) { const double2 q = pr[t >> 1]; a[t] = q.x; a[t + 1] = q.y; }
      } else {
#pragma unroll
        for (int t = 0; t < E; t += 2) { const double2 q = pr[t >> 1]; a[t] = a[t] - ci * q.x; a[t + 1] = a[t + 1] - ci * q.y; }
      }
      if (jg == g) a[kt] = (i == p) ? sinv : -ci * sinv;    // column k of the inverse-in-progress
    }
  }
  const int lds = m | 1;
  if (i < m) {
#pragma unroll
    for (int t = 0; t < E; ++t) {
      const int j = jg * E + t;
      if (j < m) sB[i + j * lds] = a[t];
    }
  }
  __syncthreads();
  // The inverse of a symmetric block is symmetric; the computed one only to cond x eps, and the Schur complement
  // S = F22 - F12^T (F11^-1 F12) is formed from the upper block alone, so that antisymmetric part would land in S, then in
  // the parent's pivot block, amplified by |W|^2 level by level (it, not the block-local pivoting, cost
  // the raw solve six digits and made structured meshes diverge).  Write back the average with the transpose.
  if (i < m) {
    const int pr = p_of[i], sr = step_of[i];
#pragma unroll
    for (int t = 0; t < E; ++t) {
      const int c = jg * E + t;
      if (c < m) F[(int64_t)c * ld + i] = 0.5 * (sB[pr + step_of[c] * lds] + sB[p_of[c] + sr * lds]);
    }
  }
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

// The same tile on the FP64 tensor path: mma.sync m8n8k4 (DMMA).  tcgen05 has no f64 kind, so warp-level DMMA is the tensor
// path Blackwell offers for this precision; on this B200 it peaks at 37.1 TFLOP/s against 34.1 for CUDA-core FMAs and 35.1
// for cuBLAS DGEMM (scripts/micro/fp64_peak.cu, profiles/r02_fp64_peak.txt), so what it buys is not peak but issue slots and
// shared-memory bandwidth: 16 MMAs (4096 FMAs) per 8 fragment loads, where the FMA tile needs 8 loads per 16 FMAs.
// 128 threads = 2 x 2 warps, each warp a 32 x 32 block of the 64 x 64 tile = 4 x 4 MMA tiles, 32 accumulators per lane.
// Fragment layout (PTX ISA, m8n8k4 .f64): A[lane >> 2][lane & 3], B[lane & 3][lane >> 2], C[lane >> 2][2 (lane & 3) + {0, 1}].
constexpr int DK = 16;        // k-slab of the DMMA tile
constexpr int DLD = GT + 4;   // leading dimension of the k-major slabs: 4 k-rows x 4 row-octets hit 16 distinct bank pairs
template <bool B_KCONTIG, bool ACCUM>
__device__ __forceinline__ void gemm_tile_dmma(const double* __restrict__ Ap, int64_t lda, const double* __restrict__ Bp,
                                               int64_t ldb, double* __restrict__ Cp, int64_t ldc, int M, int N, int K,
                                               int m0, int n0) {
  __shared__ double As[DK][DLD];
  __shared__ double Bs[DK][DLD];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = (warp & 1) * 32, wn = (warp >> 1) * 32;
  const int fr = lane >> 2, fk = lane & 3;
  double c[4][4][2];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) c[a][b][0] = c[a][b][1] = 0.0;
  for (int k0 = 0; k0 < K; k0 += DK) {
    for (int t = tid; t < GT * DK; t += 128) {
      const int kk = t % DK, i = t / DK;
      const int gi = m0 + i, gk = k0 + kk;
      As[kk][i] = (gi < M && gk < K) ? Ap[(int64_t)gi * lda + gk] : 0.0;
    }
    if (B_KCONTIG) {
      for (int t = tid; t < GT * DK; t += 128) {
        const int kk = t % DK, j = t / DK;
        const int gj = n0 + j, gk = k0 + kk;
        Bs[kk][j] = (gj < N && gk < K) ? Bp[(int64_t)gj * ldb + gk] : 0.0;
      }
    } else {
      for (int t = tid; t < GT * DK; t += 128) {
        const int j = t % GT, kk = t / GT;
        const int gj = n0 + j, gk = k0 + kk;
        Bs[kk][j] = (gj < N && gk < K) ? Bp[(int64_t)gk * ldb + gj] : 0.0;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < DK; kk += 4) {
      double av[4], bv[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) av[a] = As[kk + fk][wm + 8 * a + fr];
#pragma unroll
      for (int b = 0; b < 4; ++b) bv[b] = Bs[kk + fk][wn + 8 * b + fr];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b)
          asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                       : "+d"(c[a][b][0]), "+d"(c[a][b][1])
                       : "d"(av[a]), "d"(bv[b]));
    }
    __syncthreads();
  }
#pragma unroll
  for (int b = 0; b < 4; ++b)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int gj = n0 + wn + 8 * b + 2 * fk + e;
      if (gj >= N) continue;
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int gi = m0 + wm + 8 * a + fr;
        if (gi >= M) continue;
        double* cp = Cp + (int64_t)gj * ldc + gi;
        if (ACCUM) *cp -= c[a][b][e]; else *cp = c[a][b][e];
      }
    }
}

__global__ void __launch_bounds__(128) gemm_w_dmma_kernel(const int4* __restrict__ tiles, PlanView P) {
  const int4 t = tiles[blockIdx.x];
  const int f = t.x;
  const int s2 = 2 * P.s[f], u2 = 2 * front_u(P, f);
  const int64_t ld = s2 + u2;
  double* F = P.pool + P.foff[f];
  gemm_tile_dmma<true, false>(F + (int64_t)s2 * ld, ld, F, ld, F + s2, ld, u2, s2, s2, t.y, t.z);
}

__global__ void __launch_bounds__(128) gemm_schur_dmma_kernel(const int4* __restrict__ tiles, PlanView P) {
  const int4 t = tiles[blockIdx.x];
  const int f = t.x;
  const int s2 = 2 * P.s[f], u2 = 2 * front_u(P, f);
  const int64_t ld = s2 + u2;
  double* F = P.pool + P.foff[f];
  gemm_tile_dmma<false, true>(F + (int64_t)s2 * ld, ld, F + s2, ld, F + (int64_t)s2 * ld + s2, ld, u2, u2, s2, t.y, t.z);
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

size_t invert_smem(int m) { return (size_t)(m | 1) * m * sizeof(double); }

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
    // fronts of a level in three size classes of their pivot block (<= 32, <= 64, <= 128 unknowns): each class has its own
    // instantiation of the register-resident inverse (256 / 256 / 512 threads)
    std::vector<int32_t> lf(P.lfront);
    D.lsplit32.assign(P.nlevels, 0); D.lsplit64.assign(P.nlevels, 0);
    for (int l = 0; l < P.nlevels; ++l) {
      auto b0 = lf.begin() + P.lptr[l], e0 = lf.begin() + P.lptr[l + 1];
      auto m32 = std::stable_partition(b0, e0, [&](int32_t f) { return 2 * P.s[f] <= 32; });
      auto m64 = std::stable_partition(m32, e0, [&](int32_t f) { return 2 * P.s[f] <= 64; });
      D.lsplit32[l] = (int32_t)(m32 - b0); D.lsplit64[l] = (int32_t)(m64 - b0);
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

  std::vector<int4> wt, stl, ea;
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
    }
    D.w_ptr[l + 1] = (int32_t)wt.size(); D.s_ptr[l + 1] = (int32_t)stl.size(); D.ea_ptr[l + 1] = (int32_t)ea.size();
  }
  build_level_plan(ctx, P, uoff, in_sub, goff, D.st);
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
    PLFEM_CUDA(cudaFuncSetAttribute(invert_kernel<128, 4, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)invert_smem(MAX_PIV)));
    if (ctx->device < 64) attr_set[ctx->device] = true;
  }
  const PlanView v = view(D);
  // frontal GEMMs: FP64 tensor path (DMMA) unless PLFEM_GEMM=fma asks for the CUDA-core tile
  static const bool dmma = [] { const char* e = std::getenv("PLFEM_GEMM"); return !(e && e[0] == 'f'); }();
  for (int l = 0; l < D.nlevels; ++l) {
    const int nea = D.ea_ptr[l + 1] - D.ea_ptr[l];
    if (nea > 0) {
      extend_add_kernel<<<nea, 256, 0, ctx->stream>>>(D.ea_slabs.p + D.ea_ptr[l], v);
      ctx->launches++;
    }
    const int nfl = D.lptr[l + 1] - D.lptr[l];
    const int n32 = D.lsplit32[l], n64 = D.lsplit64[l] - D.lsplit32[l], n128 = nfl - D.lsplit64[l];
    const int32_t* lf = D.lfront.p + D.lptr[l];
    if (n32 > 0) { invert_kernel<32, 8, 4><<<n32, 256, invert_smem(32), ctx->stream>>>(lf, v, D.status.p); ctx->launches++; }
    if (n64 > 0) { invert_kernel<64, 4, 16><<<n64, 256, invert_smem(64), ctx->stream>>>(lf + n32, v, D.status.p); ctx->launches++; }
    if (n128 > 0) { invert_kernel<128, 4, 32><<<n128, 512, invert_smem(128), ctx->stream>>>(lf + n32 + n64, v, D.status.p); ctx->launches++; }
    const int nw = D.w_ptr[l + 1] - D.w_ptr[l];
    if (nw > 0) {
      if (dmma) gemm_w_dmma_kernel<<<nw, 128, 0, ctx->stream>>>(D.w_tiles.p + D.w_ptr[l], v);
      else gemm_w_kernel<<<nw, 256, 0, ctx->stream>>>(D.w_tiles.p + D.w_ptr[l], v);
      ctx->launches++;
    }
    const int ns = D.s_ptr[l + 1] - D.s_ptr[l];
    if (ns > 0) {
      if (dmma) gemm_schur_dmma_kernel<<<ns, 128, 0, ctx->stream>>>(D.s_tiles.p + D.s_ptr[l], v);
      else gemm_schur_kernel<<<ns, 256, 0, ctx->stream>>>(D.s_tiles.p + D.s_ptr[l], v);
      ctx->launches++;
    }
  }
  launch_stream_pack(ctx, D);       // the solve phase reads only the streams written here; the front pool is workspace
  PLFEM_CUDA(cudaGetLastError());
}

bool use_pdl() {
  static const bool on = [] { const char* e = std::getenv("PLFEM_PDL"); return !(e && e[0] == '0'); }();
  return on;
}

// Above the bottom subtrees a sweep is either ONE dataflow launch (persistent CTAs taking tickets, waiting on per-front
// counters: fastest for a sweep that has the device to itself, 0.195 vs 0.227 ms on a config-1 forest) or one launch per
// elimination-tree level (nothing ever waits on the device: with several forests in flight on one GPU their launches
// interleave better, 502 vs 456 solves/s).  A context says which it wants (plfem_ctx_set_sweep_schedule: the forest pool
// asks for per-level launches when it runs more than one worker); $PLFEM_SWEEP=levels|dataflow overrides for A/B runs.
bool use_fused_sweeps(const plfem_ctx* ctx) {
  static const int env = [] { const char* e = std::getenv("PLFEM_SWEEP"); return !e || !e[0] ? -1 : (e[0] == 'l' ? 1 : 0); }();
  const int mode = env >= 0 ? env : (ctx->sweep_schedule >= 0 ? ctx->sweep_schedule : 0);
  return mode == 0;
}

// nrhs right-hand sides (1 or SOLVE_NRHS), INTERLEAVED: entry i of right-hand side r at b[i * nrhs + r].  One launch for the
// bottom subtrees, then ONE dataflow launch for everything above them (sweep_stream.cu): its tasks take tickets in level
// order and wait on per-front counters, so a front starts as soon as its own children are done and no launch boundary
// (12 us each, fifteen of them per sweep of a 7-core design) separates the levels.
void run_solve_forward(plfem_ctx* ctx, const DevPlan& D, const double* b, double* z, int nrhs, const uint8_t* active) {
  if (nrhs != 1 && nrhs != SOLVE_NRHS) throw StatusError(PLFEM_ERR_INTERNAL, "unsupported number of right-hand sides");
  const bool pdl = use_pdl();
  const bool fused = use_fused_sweeps(ctx);
  reset_sweep_counters(ctx, D);                // forward and backward counters, before the first launch of the solve
  bool first = true;     // the first launch follows kernels that are not PDL-aware: plain launch
  if (D.st.n_subs > 0) {
    launch_stream_forward(ctx, D, b, z, nrhs, false, active);
    first = false;
  }
  if (fused) {
    launch_fused_forward(ctx, D, b, z, nrhs, pdl && !first, active);
    return;
  }
  for (int l = 0; l < D.nlevels; ++l) {
    if (D.st.fptr[l + 1] == D.st.fptr[l]) continue;
    launch_level_forward(ctx, D, l, b, z, nrhs, pdl && !first, active);
    first = false;
  }
}

// must follow run_solve_forward on the same stream (its launches may be PDL-chained to the forward ones); reset_counters =
// false when that forward sweep has just cleared the dataflow counters (run_solve)
void run_solve_backward(plfem_ctx* ctx, const DevPlan& D, double* x, int nrhs, bool reset_counters, const uint8_t* active) {
  const bool pdl = use_pdl();
  if (reset_counters) reset_sweep_counters(ctx, D);
  if (use_fused_sweeps(ctx)) {
    launch_fused_backward(ctx, D, x, nrhs, pdl && !reset_counters, active);
  } else {
    for (int l = D.nlevels - 1; l >= 0; --l) {
      if (D.st.bptr[l] == D.st.bptr[l + 1]) continue;
      launch_level_backward(ctx, D, l, x, nrhs, pdl && !reset_counters, active);
      reset_counters = false;
    }
  }
  if (D.st.n_subs > 0) launch_stream_backward(ctx, D, x, nrhs, pdl, active);
}

void run_solve(plfem_ctx* ctx, const DevPlan& D, const double* b, double* x, int nrhs, const uint8_t* active) {
  run_solve_forward(ctx, D, b, x, nrhs, active);
  run_solve_backward(ctx, D, x, nrhs, false, active);
}

}  // namespace plfem
