// K1/K2: element geometry + fused gather assembly of the scalar P2 matrices, and the CSR SpMV.
//
// Replaces `asm(form, basis)` x 9 and the sparse block adds of solver_fem.py:131-167.
//
// Design (B200): the classic "element kernel -> COO -> sort/scatter" pipeline writes and re-reads
// 36*T*9 triplets (2.3 KB per element) that are not algorithmic traffic.  Here one thread owns one
// structural non-zero (row node r, column node c).  It walks the elements around r in ascending
// element id (fixed summation order = deterministic, no atomics), finds c in each of them, and
// integrates the 10 scalar forms on the fly from an 96-byte per-element record (inverse Jacobian,
// |det J| and 1/eps at the 6 quadrature points) that lives in L1/L2.  DRAM traffic is the mesh in and
// each assembled value out, once.
//
// Arithmetic follows the oracle operation for operation (no FMA contraction: every product and
// sum below is an explicit round-to-nearest intrinsic), so an element-level value is exactly zero
// here iff it is exactly zero there — that decides the CSR structure scikit-fem/SciPy produce.
#include <cub/device/device_scan.cuh>

#include "common.h"

namespace plfem {

namespace {

__host__ __device__ inline double mul(double a, double b) {
#ifdef __CUDA_ARCH__
  return __dmul_rn(a, b);
#else
  volatile double r = a * b; return r;
#endif
}
__host__ __device__ inline double add(double a, double b) {
#ifdef __CUDA_ARCH__
  return __dadd_rn(a, b);
#else
  volatile double r = a + b; return r;
#endif
}

RefTables make_tables() {
  // 6-point degree-4 rule on the reference triangle and the P2 shape functions evaluated there,
  // with the same expression order as the oracle's NumPy code.
  RefTables t;
  const double a = 0.445948490915965, b = 0.091576213509771;
  const double wa = 0.111690794839005, wb = 0.054975871827661;
  const double a2 = 1 - 2 * a, b2 = 1 - 2 * b;
  const double X[6] = {a, a, a2, b, b, b2};
  const double Y[6] = {a, a2, a, b, b2, b};
  const double W[6] = {wa, wa, wa, wb, wb, wb};
  for (int q = 0; q < 6; ++q) {
    const double x = X[q], y = Y[q];
    t.qx[q] = x; t.qy[q] = y; t.w[q] = W[q];
    const double xx = mul(x, x), yy = mul(y, y), xy4 = mul(mul(4.0, x), y);
    // phi
    t.phi[0 * 6 + q] = add(add(add(add(add(1.0, -mul(3.0, x)), -mul(3.0, y)), mul(2.0, xx)), xy4), mul(2.0, yy));
    t.phi[1 * 6 + q] = add(mul(2.0, xx), -x);
    t.phi[2 * 6 + q] = add(mul(2.0, yy), -y);
    t.phi[3 * 6 + q] = add(add(mul(4.0, x), -mul(4.0, xx)), -xy4);
    t.phi[4 * 6 + q] = xy4;
    t.phi[5 * 6 + q] = add(add(mul(4.0, y), -xy4), -mul(4.0, yy));
    // d/dx
    t.dx[0 * 6 + q] = add(add(-3.0, mul(4.0, x)), mul(4.0, y));
    t.dx[1 * 6 + q] = add(mul(4.0, x), -1.0);
    t.dx[2 * 6 + q] = mul(0.0, x);
    t.dx[3 * 6 + q] = add(add(4.0, -mul(8.0, x)), -mul(4.0, y));
    t.dx[4 * 6 + q] = mul(4.0, y);
    t.dx[5 * 6 + q] = mul(-4.0, y);
    // d/dy
    t.dy[0 * 6 + q] = add(add(-3.0, mul(4.0, x)), mul(4.0, y));
    t.dy[1 * 6 + q] = mul(0.0, x);
    t.dy[2 * 6 + q] = add(mul(4.0, y), -1.0);
    t.dy[3 * 6 + q] = mul(-4.0, x);
    t.dy[4 * 6 + q] = mul(4.0, x);
    t.dy[5 * 6 + q] = add(add(4.0, -mul(4.0, x)), -mul(8.0, y));
  }
  return t;
}

__constant__ RefTables c_tab;

constexpr int ELEM_STRIDE = 12;  // doubles per element record: i00 i10 i01 i11 |det| w[6] pad

// ---- K1a: one thread per element -------------------------------------------------------------------
__global__ void element_setup_kernel(const double* __restrict__ px, const double* __restrict__ py,
                                     const int32_t* __restrict__ edofs, int64_t T, const double* __restrict__ cores,
                                     int ncores, double eps_core, double eps_clad,
                                     const double* __restrict__ eps_at_quad, bool store_eps, double* __restrict__ elem) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= T) return;
  const int32_t v0 = edofs[6 * e], v1 = edofs[6 * e + 1], v2 = edofs[6 * e + 2];
  const double x0 = px[v0], y0 = py[v0];
  const double a00 = add(px[v1], -x0), a01 = add(px[v2], -x0);
  const double a10 = add(py[v1], -y0), a11 = add(py[v2], -y0);
  const double det = add(mul(a00, a11), -mul(a01, a10));
  double* r = elem + e * ELEM_STRIDE;
  r[0] = __ddiv_rn(a11, det);    // invA[0][0]
  r[1] = __ddiv_rn(-a10, det);   // invA[1][0]
  r[2] = __ddiv_rn(-a01, det);   // invA[0][1]
  r[3] = __ddiv_rn(a00, det);    // invA[1][1]
  r[4] = fabs(det);
  for (int q = 0; q < 6; ++q) {
    double eps;
    if (eps_at_quad) {
      eps = eps_at_quad[e * 6 + q];
    } else {
      const double xq = add(add(mul(a00, c_tab.qx[q]), mul(a01, c_tab.qy[q])), x0);
      const double yq = add(add(mul(a10, c_tab.qx[q]), mul(a11, c_tab.qy[q])), y0);
      eps = eps_clad;
      for (int c = 0; c < ncores; ++c) {
        const double ddx = add(xq, -cores[3 * c]), ddy = add(yq, -cores[3 * c + 1]), rr = cores[3 * c + 2];
        if (add(mul(ddx, ddx), mul(ddy, ddy)) <= mul(rr, rr)) eps = eps_core;
      }
    }
    r[5 + q] = store_eps ? eps : __ddiv_rn(1.0, eps);     // weight of the "w" forms: 1/eps (H-field system) or eps (scalar Helmholtz)
  }
  r[11] = 0.0;
}

__global__ void expand_rows_kernel(const int32_t* __restrict__ rowptr, int32_t n, int32_t* __restrict__ rowidx) {
  const int32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  for (int32_t z = rowptr[r]; z < rowptr[r + 1]; ++z) rowidx[z] = r;
}


// ---- sparsity pattern of the interior nodes in elimination order, built on the device ---------------------------
// Row R (a node in elimination order) holds the nodes that share an element with it — the union over the elements
// around its DOF of their 6 DOFs, dropped when eliminated by the Dirichlet condition, sorted ascending.  One WARP per
// row: the 6 x (elements around the node) candidate columns go to shared memory, duplicates are struck out (a candidate
// survives if no earlier one has its value), and a survivor's position in the row is the number of smaller survivors —
// a few dozen broadcast reads per lane, no per-thread list.  COUNT pass -> inclusive scan -> FILL pass.  Identical to the
// host `build_pattern` (symbolic.cpp), which the export path still uses; here it saves ~5 ms of host time per design,
// which is what a forest pool is short of.
constexpr int PATTERN_MAXCAND = 768;     // 6 x 128 elements around one node (a core centre is the hub of a whole ring)
constexpr int PATTERN_WARPS = 8;

template <bool FILL>
__global__ void __launch_bounds__(32 * PATTERN_WARPS)
pattern_rows_kernel(int32_t r0, int32_t n, const int32_t* __restrict__ old_of_new, const int32_t* __restrict__ n2e_ptr,
                    const int32_t* __restrict__ n2e, const int32_t* __restrict__ edofs, const int32_t* __restrict__ new_of_dof,
                    int32_t* __restrict__ rowptr, int32_t* __restrict__ col, int32_t* __restrict__ rowidx, int32_t* __restrict__ err) {
  __shared__ int32_t s_cand[PATTERN_WARPS][PATTERN_MAXCAND];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int32_t r = blockIdx.x * PATTERN_WARPS + w;
  if (r >= n) return;                                   // whole warps leave together
  int32_t* cand = s_cand[w];
  const int32_t R = r0 + r;
  const int32_t o = old_of_new[R];
  const int32_t q0 = n2e_ptr[o];
  int nc = 6 * (n2e_ptr[o + 1] - q0);
  if (nc > PATTERN_MAXCAND) { if (lane == 0) atomicExch(err, 1); nc = PATTERN_MAXCAND; }
  for (int i = lane; i < nc; i += 32) cand[i] = new_of_dof[edofs[6 * (int64_t)n2e[q0 + i / 6] + i % 6]];
  __syncwarp();
  // strike out duplicates: decided from the untouched list, written after everyone has read (bit t of `dup` = candidate
  // lane + 32 t is a repetition of an earlier one)
  uint32_t dup = 0;
  for (int t = 0; 32 * t < nc; ++t) {
    const int i = lane + 32 * t;
    if (i < nc) {
      const int32_t c = cand[i];
      bool d = false;
      for (int j = 0; j < i; ++j) d |= (cand[j] == c);
      dup |= (uint32_t)d << t;
    }
  }
  __syncwarp();
  for (int t = 0; 32 * t < nc; ++t) {
    const int i = lane + 32 * t;
    if (i < nc && ((dup >> t) & 1u)) cand[i] = -1;
  }
  __syncwarp();
  int mine = 0;
  const int32_t z0 = FILL ? rowptr[R] : 0;
  for (int t = 0; 32 * t < nc; ++t) {
    const int i = lane + 32 * t;
    if (i >= nc) continue;
    const int32_t c = cand[i];
    if (c < 0) continue;                                  // eliminated by the Dirichlet condition, or a duplicate
    ++mine;
    if (FILL) {
      int rank = 0;
      for (int j = 0; j < nc; ++j) { const int32_t cj = cand[j]; rank += (cj >= 0 && cj < c); }
      col[z0 + rank] = c;
      rowidx[z0 + rank] = R;
    }
  }
  if (!FILL) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) mine += __shfl_down_sync(0xffffffffu, mine, off);
    if (lane == 0) rowptr[R + 1] = mine;
  }
}

__global__ void scatter_new_of_dof_kernel(int32_t r0, int32_t n, const int32_t* __restrict__ old_of_new, int32_t* __restrict__ new_of_dof) {
  const int32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n) new_of_dof[old_of_new[r0 + r]] = r0 + r;
}

__global__ void gather_offsets_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ node_off, int nb1, int32_t* __restrict__ out) {
  const int i = threadIdx.x;
  if (i < nb1) out[i] = rowptr[node_off[i]];
}

// ---- K1b/K2 fused: quadrature per (row node, element) pair, combined per structural non-zero inside the CTA ---------------
// MODE 0: the solve's value arrays of the H-field system; 1: the ten scalar matrices for export; 2: scalar Helmholtz pencil
// (solver_fem.py:245-276) in the same arrays — Hx block (K - k0^2 M_eps, M), Hy block (shift M, M), no coupling.
//
// Round 1 gave every structural non-zero (r, c) a thread that walked the elements around r looking for c: 0.044 of the HBM
// peak, because the thread of a diagonal entry integrates over ALL elements around its node (6 for a vertex) while its
// warp-mates integrate over one or two, and every thread repeats the search.  Now a CTA takes ASM_RPB consecutive pattern rows:
//   phase 1  one thread per (row node r, element e around r): the test-function side (gradients and value of r's shape function
//            at the 6 quadrature points) is computed once, then the 10 forms against each of the element's 6 trial functions —
//            36 T uniform work items per mesh, no search, no divergence; the 6 x 10 element-level values go to shared memory;
//   phase 2  one thread per structural non-zero of those rows: adds the staged values of its column over the row's elements in
//            ascending element id (the same fixed order, operation for operation, as the walk: bit-identical results), applies
//            the block formulas and writes each assembled value once, coalesced.
// DRAM traffic stays the mesh in and every value out once; the element-level values never leave the SM.  The CTA's rows are
// taken in batches of at most ASM_CAP pairs (vertex nodes cluster in the elimination order: 6 elements each against 2 for an
// edge node, so a fixed row count per batch either overflows the staging area or leaves the threads idle); a single row with
// more than ASM_CAP elements (a hub node of a ring of 100+ elements) takes the old per-entry walk.
constexpr int ASM_T = 128;       // threads per CTA
constexpr int ASM_RPB = 80;      // pattern rows per CTA: 3 elements around a P2 node on average -> two batches of <= ASM_CAP pairs
constexpr int ASM_CAP = 128;     // (row, element) pairs staged per CTA (60 doubles each: 61 KB)

struct AsmArgs {
  int32_t row0, nrows;                               // this design's rows of the (possibly concatenated) pattern
  const int32_t *rowptr, *col, *old_of_new;          // pattern (global row / non-zero ids)
  const int32_t *n2e_ptr, *n2e, *edofs;              // this design's mesh tables
  const double* elem;
  double k0sq, alpha;
  double* vals; int64_t vstride; uint32_t* flags;
};

struct AsmSmem {
  double l[60 * ASM_CAP];        // [(j * 10 + k) * ASM_CAP + pair]
  int32_t key[6 * ASM_CAP];      // DOF id of trial function j of the pair's element
  double phi[36], dx[36], dy[36], w[6];
  int32_t o[ASM_RPB], q0[ASM_RPB], pp[ASM_RPB + 1], rp[ASM_RPB + 1];
};

// the 10 element-level forms of (test i, trial j) of one element, oracle operation order
__device__ __forceinline__ void element_forms(const AsmSmem& sm, const double* __restrict__ r, int li, int lj, double (&l)[10]) {
  const double i00 = r[0], i10 = r[1], i01 = r[2], i11 = r[3], adet = r[4];
#pragma unroll
  for (int k = 0; k < 10; ++k) l[k] = 0.0;
#pragma unroll
  for (int qp = 0; qp < 6; ++qp) {
    const double w = r[5 + qp];
    const double dxq = mul(adet, sm.w[qp]);
    // test function v = row node (i), trial function u = column node (j)
    const double gxi = add(mul(i00, sm.dx[li * 6 + qp]), mul(i10, sm.dy[li * 6 + qp]));
    const double gyi = add(mul(i01, sm.dx[li * 6 + qp]), mul(i11, sm.dy[li * 6 + qp]));
    const double gxj = add(mul(i00, sm.dx[lj * 6 + qp]), mul(i10, sm.dy[lj * 6 + qp]));
    const double gyj = add(mul(i01, sm.dx[lj * 6 + qp]), mul(i11, sm.dy[lj * 6 + qp]));
    const double pi = sm.phi[li * 6 + qp], pj = sm.phi[lj * 6 + qp];
    l[X_KXX] = add(l[X_KXX], mul(mul(mul(w, gyj), gyi), dxq));
    l[X_KYY] = add(l[X_KYY], mul(mul(mul(w, gxj), gxi), dxq));
    l[X_KXY] = add(l[X_KXY], mul(mul(mul(-w, gyj), gxi), dxq));
    l[X_KYX] = add(l[X_KYX], mul(mul(mul(-w, gxj), gyi), dxq));
    l[X_DXX] = add(l[X_DXX], mul(mul(gxj, gxi), dxq));
    l[X_DYY] = add(l[X_DYY], mul(mul(gyj, gyi), dxq));
    l[X_DXY] = add(l[X_DXY], mul(mul(gxj, gyi), dxq));
    l[X_DYX] = add(l[X_DYX], mul(mul(gxi, gyj), dxq));   // Dxy(c, r): trial = row node, test = column node
    l[X_M] = add(l[X_M], mul(mul(pj, pi), dxq));
    l[X_MINV] = add(l[X_MINV], mul(mul(mul(w, pj), pi), dxq));
  }
}

template <int MODE>
__device__ __forceinline__ void store_entry(const AsmArgs& a, int64_t z, const double (&g)[10], uint32_t fl) {
  double* vals = a.vals; const int64_t vstride = a.vstride; const double k0sq = a.k0sq, alpha = a.alpha;
  if (MODE == 1) {
#pragma unroll
    for (int k = 0; k < 10; ++k) vals[(int64_t)k * vstride + z] = g[k];
    a.flags[z] = fl;
  } else if (MODE == 2) {
    // the element records hold eps (not 1/eps) in this mode: g[X_MINV] is M_eps; `alpha` carries the decoupled shift
    vals[(int64_t)S_AXX * vstride + z] = add(add(g[X_DXX], g[X_DYY]), -mul(k0sq, g[X_MINV]));
    vals[(int64_t)S_AXY * vstride + z] = 0.0;
    vals[(int64_t)S_AYX * vstride + z] = 0.0;
    vals[(int64_t)S_AYY * vstride + z] = mul(alpha, g[X_M]);
    vals[(int64_t)S_MINV * vstride + z] = g[X_M];
    vals[(int64_t)S_DXX * vstride + z] = 0.0;
    vals[(int64_t)S_DXY * vstride + z] = 0.0;
    vals[(int64_t)S_DYY * vstride + z] = 0.0;
  } else {
    const double km = mul(k0sq, g[X_M]);
    vals[(int64_t)S_AXX * vstride + z] = add(add(g[X_KXX], mul(alpha, g[X_DXX])), -km);
    vals[(int64_t)S_AXY * vstride + z] = add(g[X_KXY], mul(alpha, g[X_DXY]));
    vals[(int64_t)S_AYX * vstride + z] = add(g[X_KYX], mul(alpha, g[X_DYX]));
    vals[(int64_t)S_AYY * vstride + z] = add(add(g[X_KYY], mul(alpha, g[X_DYY])), -km);
    vals[(int64_t)S_MINV * vstride + z] = g[X_MINV];
    vals[(int64_t)S_DXX * vstride + z] = g[X_DXX];
    vals[(int64_t)S_DXY * vstride + z] = g[X_DXY];
    vals[(int64_t)S_DYY * vstride + z] = g[X_DYY];
  }
}

// where pair q keeps its value (trial function lj, form k): rotated by 3 lj so that the six columns of one element, read by six
// threads of phase 2 in the same cycle, sit in six different banks; a warp of phase 1 still writes 32 consecutive doubles
__device__ __forceinline__ int stage_at(int lj, int k, int q) { return (lj * 10 + k) * ASM_CAP + ((q + 3 * lj) & (ASM_CAP - 1)); }

// per-entry walk over the elements around the row node (the round-1 scheme): only for a row with more than ASM_CAP elements
template <int MODE>
__device__ __forceinline__ void walk_entry(const AsmArgs& a, const AsmSmem& sm, int32_t orow, int32_t z) {
  const int32_t ocol = a.old_of_new[a.col[z]];
  double g[10];
#pragma unroll
  for (int k = 0; k < 10; ++k) g[k] = 0.0;
  uint32_t fl = 0;
  for (int32_t q = a.n2e_ptr[orow]; q < a.n2e_ptr[orow + 1]; ++q) {
    const int32_t e = a.n2e[q];
    const int32_t* ed = a.edofs + 6 * (int64_t)e;
    int li = -1, lj = -1;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      const int32_t d = ed[k];
      if (d == orow) li = k;
      if (d == ocol) lj = k;
    }
    if (lj < 0) continue;
    double l[10];
    element_forms(sm, a.elem + (int64_t)e * ELEM_STRIDE, li, lj, l);
#pragma unroll
    for (int k = 0; k < 10; ++k) {
      g[k] = add(g[k], l[k]);
      if (MODE == 1 && l[k] != 0.0) fl |= (1u << k);   // NaN != 0 is true: a NaN entry is stored, as in SciPy
    }
  }
  store_entry<MODE>(a, z, g, fl);
}

template <int MODE>
__global__ void __launch_bounds__(ASM_T) assemble_rows_kernel(AsmArgs a) {
  extern __shared__ __align__(16) unsigned char asm_smem_raw[];
  AsmSmem& sm = *reinterpret_cast<AsmSmem*>(asm_smem_raw);
  const int tid = threadIdx.x;
  for (int i = tid; i < 36; i += ASM_T) { sm.phi[i] = c_tab.phi[i]; sm.dx[i] = c_tab.dx[i]; sm.dy[i] = c_tab.dy[i]; }
  if (tid < 6) sm.w[tid] = c_tab.w[tid];
  const int32_t R0 = a.row0 + blockIdx.x * ASM_RPB;
  const int nr = min(ASM_RPB, a.row0 + a.nrows - R0);
  if (tid < nr) {
    const int32_t o = a.old_of_new[R0 + tid];
    sm.o[tid] = o; sm.q0[tid] = a.n2e_ptr[o];
    sm.pp[tid + 1] = a.n2e_ptr[o + 1] - a.n2e_ptr[o];
  }
  if (tid <= nr) sm.rp[tid] = a.rowptr[R0 + tid];
  __syncthreads();
  if (tid == 0) {
    int acc = 0; sm.pp[0] = 0;
    for (int i = 0; i < nr; ++i) { acc += sm.pp[i + 1]; sm.pp[i + 1] = acc; }
  }
  __syncthreads();
  // the CTA's rows in batches of at most ASM_CAP (row, element) pairs: every thread derives the same batches
  for (int rb = 0; rb < nr;) {
    int re = rb + 1;
    const int p0 = sm.pp[rb];
    while (re < nr && sm.pp[re + 1] - p0 <= ASM_CAP) ++re;
    const int npairs = sm.pp[re] - p0;
    const int32_t zb = sm.rp[rb], ze = sm.rp[re];
    if (npairs > ASM_CAP) {       // one hub row on its own
      for (int32_t z = zb + tid; z < ze; z += ASM_T) walk_entry<MODE>(a, sm, sm.o[rb], z);
      rb = re;
      continue;
    }
    // phase 1: one (row node, element) pair per thread
    for (int q = tid; q < npairs; q += ASM_T) {
      const int p = p0 + q;
      int lo = rb, hi = re;
      while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (sm.pp[mid] <= p) lo = mid; else hi = mid; }
      const int32_t orow = sm.o[lo];
      const int32_t e = a.n2e[sm.q0[lo] + (p - sm.pp[lo])];
      const int32_t* ed = a.edofs + 6 * (int64_t)e;
      int li = 0;
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        const int32_t d = ed[k];
        sm.key[k * ASM_CAP + q] = d;
        if (d == orow) li = k;
      }
      const double* r = a.elem + (int64_t)e * ELEM_STRIDE;
      const double i00 = r[0], i10 = r[1], i01 = r[2], i11 = r[3], adet = r[4];
      // test-function side and weights at the 6 quadrature points: once per pair (same operations as element_forms)
      double gxi[6], gyi[6], pi[6], wq[6], dxq[6];
#pragma unroll
      for (int qp = 0; qp < 6; ++qp) {
        wq[qp] = r[5 + qp];
        dxq[qp] = mul(adet, sm.w[qp]);
        gxi[qp] = add(mul(i00, sm.dx[li * 6 + qp]), mul(i10, sm.dy[li * 6 + qp]));
        gyi[qp] = add(mul(i01, sm.dx[li * 6 + qp]), mul(i11, sm.dy[li * 6 + qp]));
        pi[qp] = sm.phi[li * 6 + qp];
      }
#pragma unroll 1
      for (int lj = 0; lj < 6; ++lj) {
        double l[10];
#pragma unroll
        for (int k = 0; k < 10; ++k) l[k] = 0.0;
#pragma unroll
        for (int qp = 0; qp < 6; ++qp) {
          const double w = wq[qp];
          const double gxj = add(mul(i00, sm.dx[lj * 6 + qp]), mul(i10, sm.dy[lj * 6 + qp]));
          const double gyj = add(mul(i01, sm.dx[lj * 6 + qp]), mul(i11, sm.dy[lj * 6 + qp]));
          const double pj = sm.phi[lj * 6 + qp];
          const double wgy = mul(w, gyj), wgx = mul(w, gxj);       // mul(-w, g) == -mul(w, g) exactly
          l[X_KXX] = add(l[X_KXX], mul(mul(wgy, gyi[qp]), dxq[qp]));
          l[X_KYY] = add(l[X_KYY], mul(mul(wgx, gxi[qp]), dxq[qp]));
          l[X_KXY] = add(l[X_KXY], mul(mul(-wgy, gxi[qp]), dxq[qp]));
          l[X_KYX] = add(l[X_KYX], mul(mul(-wgx, gyi[qp]), dxq[qp]));
          l[X_DXX] = add(l[X_DXX], mul(mul(gxj, gxi[qp]), dxq[qp]));
          l[X_DYY] = add(l[X_DYY], mul(mul(gyj, gyi[qp]), dxq[qp]));
          l[X_DXY] = add(l[X_DXY], mul(mul(gxj, gyi[qp]), dxq[qp]));
          l[X_DYX] = add(l[X_DYX], mul(mul(gxi[qp], gyj), dxq[qp]));
          l[X_M] = add(l[X_M], mul(mul(pj, pi[qp]), dxq[qp]));
          l[X_MINV] = add(l[X_MINV], mul(mul(mul(w, pj), pi[qp]), dxq[qp]));
        }
#pragma unroll
        for (int k = 0; k < 10; ++k) sm.l[stage_at(lj, k, q)] = l[k];
      }
    }
    __syncthreads();
    // phase 2: one structural non-zero per thread
    for (int32_t z = zb + tid; z < ze; z += ASM_T) {
      int lo = rb, hi = re;
      while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (sm.rp[mid] <= z) lo = mid; else hi = mid; }
      const int32_t ocol = a.old_of_new[a.col[z]];
      double g[10];
#pragma unroll
      for (int k = 0; k < 10; ++k) g[k] = 0.0;
      uint32_t fl = 0;
      for (int q = sm.pp[lo] - p0; q < sm.pp[lo + 1] - p0; ++q) {      // the row's elements, ascending id
        int lj = -1;
#pragma unroll
        for (int k = 0; k < 6; ++k) lj = (sm.key[k * ASM_CAP + q] == ocol) ? k : lj;
        if (lj < 0) continue;
#pragma unroll
        for (int k = 0; k < 10; ++k) {
          const double l = sm.l[stage_at(lj, k, q)];
          g[k] = add(g[k], l);
          if (MODE == 1 && l != 0.0) fl |= (1u << k);
        }
      }
      store_entry<MODE>(a, z, g, fl);
    }
    __syncthreads();              // the staging area is reused by the next batch
    rb = re;
  }
}

// ---- K3: CSR SpMV, a group of TPR threads per row ------------------------------------------------------
template <int TPR>
__global__ void __launch_bounds__(256)
spmv_csr_kernel(int64_t rows, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                const double* __restrict__ val, const double* __restrict__ x, double* __restrict__ y) {
  const int64_t gid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t row = gid / TPR;
  const int lane = (int)(gid % TPR);
  double acc = 0.0;
  if (row < rows) {
    const int32_t b = rowptr[row], e = rowptr[row + 1];
    for (int32_t k = b + lane; k < e; k += TPR) acc = fma(val[k], __ldg(x + col[k]), acc);
  }
#pragma unroll
  for (int off = TPR / 2; off > 0; off >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, off, TPR);
  if (row < rows && lane == 0) y[row] = acc;
}

}  // namespace

const RefTables& ref_tables() {
  static const RefTables t = make_tables();
  return t;
}

// Called once per context, synchronously: contexts on other host threads may launch kernels that read
// the constant bank right after, on their own streams.
void upload_tables() {
  const RefTables& t = ref_tables();
  PLFEM_CUDA(cudaMemcpyToSymbol(c_tab, &t, sizeof(RefTables), 0, cudaMemcpyHostToDevice));
}

void launch_element_setup(plfem_ctx* ctx, const double* d_p, const int32_t* d_edofs, int64_t V, int64_t T,
                          const plfem_material& mat, const double* d_cores, const double* d_eps_at_quad,
                          double* d_elem) {
  const int bs = 128;
  element_setup_kernel<<<(unsigned)((T + bs - 1) / bs), bs, 0, ctx->stream>>>(
      d_p, d_p + V, d_edofs, T, d_cores, mat.n_cores, mat.eps_core, mat.eps_clad, d_eps_at_quad, mat.scalar_mode != 0, d_elem);
  PLFEM_CUDA(cudaGetLastError());
  ctx->launches++;
}

void launch_expand_rows(plfem_ctx* ctx, const DevPattern& pat) {
  const int bs = 256;
  expand_rows_kernel<<<(pat.n + bs - 1) / bs, bs, 0, ctx->stream>>>(pat.rowptr.p, pat.n, pat.rowidx.p);
  PLFEM_CUDA(cudaGetLastError());
  ctx->launches++;
}

void build_device_pattern(plfem_ctx* ctx, int nb, const PatternSource* src, const std::vector<int32_t>& node_off,
                          const std::vector<int32_t>& old_of_new, DevPattern& D, std::vector<int64_t>& nnz_off) {
  cudaStream_t st = ctx->stream;
  const int32_t n_tot = node_off[nb];
  D.n = n_tot;
  D.old_of_new.upload(ctx, old_of_new);
  D.rowptr.alloc(ctx, (size_t)n_tot + 1);
  int64_t dof_tot = 0;
  std::vector<int64_t> dof_off(nb + 1, 0);
  for (int b = 0; b < nb; ++b) dof_off[b + 1] = dof_off[b] + src[b].N;
  dof_tot = dof_off[nb];
  DevBuf<int32_t> new_of_dof, d_noff, d_zoff, d_err;
  new_of_dof.alloc(ctx, (size_t)dof_tot);
  PLFEM_CUDA(cudaMemsetAsync(new_of_dof.p, 0xFF, (size_t)dof_tot * sizeof(int32_t), st));      // -1: not an interior node
  PLFEM_CUDA(cudaMemsetAsync(D.rowptr.p, 0, sizeof(int32_t), st));
  d_err.alloc(ctx, 1);
  PLFEM_CUDA(cudaMemsetAsync(d_err.p, 0, sizeof(int32_t), st));
  d_noff.upload(ctx, node_off);
  d_zoff.alloc(ctx, (size_t)nb + 1);
  for (int b = 0; b < nb; ++b) {
    const int32_t r0 = node_off[b], n = node_off[b + 1] - node_off[b];
    if (n == 0) continue;
    scatter_new_of_dof_kernel<<<(n + 255) / 256, 256, 0, st>>>(r0, n, D.old_of_new.p, new_of_dof.p + dof_off[b]);
    pattern_rows_kernel<false><<<(n + PATTERN_WARPS - 1) / PATTERN_WARPS, 32 * PATTERN_WARPS, 0, st>>>(r0, n, D.old_of_new.p, src[b].n2e_ptr, src[b].n2e, src[b].edofs,
                                                                new_of_dof.p + dof_off[b], D.rowptr.p, nullptr, nullptr, d_err.p);
    ctx->launches += 2;
  }
  PLFEM_CUDA(cudaGetLastError());
  size_t tmp_bytes = 0;
  PLFEM_CUDA(cub::DeviceScan::InclusiveSum(nullptr, tmp_bytes, D.rowptr.p, D.rowptr.p, n_tot + 1, st));
  DevBuf<uint8_t> tmp;
  tmp.alloc(ctx, std::max<size_t>(tmp_bytes, 1));
  PLFEM_CUDA(cub::DeviceScan::InclusiveSum(tmp.p, tmp_bytes, D.rowptr.p, D.rowptr.p, n_tot + 1, st));
  gather_offsets_kernel<<<1, 32 * ((nb + 1 + 31) / 32), 0, st>>>(D.rowptr.p, d_noff.p, nb + 1, d_zoff.p);
  ctx->launches += 3;      // the scan is two kernels (tile-state init + scan)
  std::vector<int32_t> zoff(nb + 1), err(1);
  d_zoff.download(zoff.data(), zoff.size());
  d_err.download(err.data(), 1);
  PLFEM_CUDA(stream_wait(st));
  if (err[0]) throw StatusError(PLFEM_ERR_INVALID, "a mesh node belongs to more than " + std::to_string(PATTERN_MAXCAND / 6) + " elements");
  nnz_off.assign(nb + 1, 0);
  for (int b = 0; b <= nb; ++b) nnz_off[b] = zoff[b];
  D.nnz = nnz_off[nb];
  D.col.alloc(ctx, std::max<size_t>((size_t)D.nnz, 1));
  D.rowidx.alloc(ctx, std::max<size_t>((size_t)D.nnz, 1));
  for (int b = 0; b < nb; ++b) {
    const int32_t r0 = node_off[b], n = node_off[b + 1] - node_off[b];
    if (n == 0) continue;
    pattern_rows_kernel<true><<<(n + PATTERN_WARPS - 1) / PATTERN_WARPS, 32 * PATTERN_WARPS, 0, st>>>(r0, n, D.old_of_new.p, src[b].n2e_ptr, src[b].n2e, src[b].edofs,
                                                               new_of_dof.p + dof_off[b], D.rowptr.p, D.col.p, D.rowidx.p, d_err.p);
    ctx->launches++;
  }
  PLFEM_CUDA(cudaGetLastError());
  // the scratch buffers return to the context's arena here; it serves this stream only, so reuse is stream-ordered
}

// one design's rows [row0, row0 + nrows) of a (possibly concatenated) pattern; values written to vals[k * vstride + z] with z
// the global non-zero id; old_of_new maps pattern rows / columns to DOF ids of THIS design's mesh
void launch_assemble_slice(plfem_ctx* ctx, int32_t row0, int32_t nrows, const DevPattern& pat, const int32_t* d_n2e_ptr,
                           const int32_t* d_n2e, const int32_t* d_edofs, const double* d_elem, double k0sq, double alpha, int mode,
                           double* d_vals, int64_t vstride, uint32_t* d_flags) {
  if (nrows <= 0) return;
  static thread_local bool configured = false;
  if (!configured) {
    PLFEM_CUDA(cudaFuncSetAttribute(assemble_rows_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(AsmSmem)));
    PLFEM_CUDA(cudaFuncSetAttribute(assemble_rows_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(AsmSmem)));
    PLFEM_CUDA(cudaFuncSetAttribute(assemble_rows_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(AsmSmem)));
    configured = true;
  }
  const AsmArgs a{row0, nrows, pat.rowptr.p, pat.col.p, pat.old_of_new.p, d_n2e_ptr, d_n2e, d_edofs, d_elem, k0sq, alpha, d_vals, vstride, d_flags};
  const unsigned grid = (unsigned)((nrows + ASM_RPB - 1) / ASM_RPB);
  if (mode == 1) assemble_rows_kernel<1><<<grid, ASM_T, sizeof(AsmSmem), ctx->stream>>>(a);
  else if (mode == 2) assemble_rows_kernel<2><<<grid, ASM_T, sizeof(AsmSmem), ctx->stream>>>(a);
  else assemble_rows_kernel<0><<<grid, ASM_T, sizeof(AsmSmem), ctx->stream>>>(a);
  PLFEM_CUDA(cudaGetLastError());
  ctx->launches++;
}

void launch_assemble(plfem_ctx* ctx, const DevPattern& pat, const int32_t* d_n2e_ptr, const int32_t* d_n2e,
                     const int32_t* d_edofs, const double* d_elem, double k0sq, double alpha, int mode,
                     double* d_vals, uint32_t* d_flags) {
  launch_assemble_slice(ctx, 0, pat.n, pat, d_n2e_ptr, d_n2e, d_edofs, d_elem, k0sq, alpha, mode, d_vals, pat.nnz, d_flags);
}

void launch_spmv_csr(plfem_ctx* ctx, int64_t rows, const int32_t* rowptr, const int32_t* col, const double* val,
                     const double* x, double* y) {
  constexpr int TPR = 8;
  const int bs = 256;
  const unsigned grid = (unsigned)((rows * TPR + bs - 1) / bs);
  spmv_csr_kernel<TPR><<<grid, bs, 0, ctx->stream>>>(rows, rowptr, col, val, x, y);
  PLFEM_CUDA(cudaGetLastError());
  ctx->launches++;
}

}  // namespace plfem
