// K7-K8: shift-invert thick-restart Lanczos in the B inner product, and the per-mode reductions.
//
// Replaces scipy eigsh(A_int, k, M=B_int, sigma, which='LM', tol, maxiter) = ARPACK dsaupd mode 3
// (solver_fem.py:197) and the per-mode loop of solver_fem.py:206-220.
//
// OP = (A - sigma B)^-1 B is self-adjoint in <x,y>_B.  The Krylov basis V and its image BV stay in
// HBM; every scalar of the recurrence (alpha, beta, the CGS2 coefficients) stays on the device too,
// so a run of Lanczos steps is a pure stream of launches with no host synchronisation.  The host is
// only involved at a restart: it reads back <= 2*ncv scalars, diagonalises the ncv x ncv projected
// matrix (symeig.cpp) and sends the rotation back; the rotation itself (V <- V S) is a device kernel.
// Thick restart with exact shifts is mathematically what ARPACK's implicit restart does.
#include "common.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <atomic>
#include <exception>
#include <mutex>
#include <numeric>
#include <string>
#include <thread>

namespace plfem {

namespace {

constexpr int RED_T = 1024;

__device__ __forceinline__ double block_sum(double v, double* sh) {
  // deterministic CTA reduction (fixed tree), blockDim.x == RED_T or smaller multiple of 32
  const int lane = threadIdx.x % 32, warp = threadIdx.x / 32;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  const int nw = blockDim.x / 32;
  v = (threadIdx.x < nw) ? sh[threadIdx.x] : 0.0;
  if (warp == 0) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
  }
  return v;  // valid in thread 0
}

// Y = B X on the permuted interleaved layout, NR right-hand sides (columns of ld2 double2):
// y[r] (2 comps) = sum_z minv[z] * x[col[z]] (2 comps)
template <int NR>
__global__ void __launch_bounds__(256) spmm_b_kernel(int32_t n, const int32_t* __restrict__ rowptr,
                                                     const int32_t* __restrict__ col, const double* __restrict__ minv,
                                                     const double2* __restrict__ x, double2* __restrict__ y, int64_t ld2) {
  constexpr int TPR = 4;
  const int64_t gid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t row = gid / TPR;
  const int lane = (int)(gid % TPR);
  double ax[NR], ay[NR];
#pragma unroll
  for (int r = 0; r < NR; ++r) { ax[r] = 0.0; ay[r] = 0.0; }
  if (row < n) {
    for (int32_t z = rowptr[row] + lane; z < rowptr[row + 1]; z += TPR) {
      const double m = minv[z];
      const int32_t c = col[z];
#pragma unroll
      for (int r = 0; r < NR; ++r) {
        const double2 v = x[r * ld2 + c];
        ax[r] = fma(m, v.x, ax[r]); ay[r] = fma(m, v.y, ay[r]);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < NR; ++r) {
#pragma unroll
    for (int off = TPR / 2; off > 0; off >>= 1) {
      ax[r] += __shfl_down_sync(0xffffffffu, ax[r], off, TPR);
      ay[r] += __shfl_down_sync(0xffffffffu, ay[r], off, TPR);
    }
    if (row < n && lane == 0) y[r * ld2 + row] = make_double2(ax[r], ay[r]);
  }
}

// T = Bv - (A - sigma B) X (iterative refinement of the block-LDL^T solve).  X, Bv and T hold NR interleaved right-hand
// sides in the permuted layout: unknown i of right-hand side r at x[i*NR + r], i.e. the 2*NR values of a node are one
// contiguous 16*NR-byte record
template <int NR>
__global__ void __launch_bounds__(256) resid_k_kernel(int32_t n, const int32_t* __restrict__ rowptr,
                                                      const int32_t* __restrict__ col, const double* __restrict__ vals,
                                                      int64_t nnz, const double* __restrict__ sigma_node,
                                                      const double* __restrict__ x, const double* __restrict__ b,
                                                      double* __restrict__ t, const uint8_t* __restrict__ active) {
  constexpr int TPR = 4;
  const int64_t gid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t row = gid / TPR;
  const int lane = (int)(gid % TPR);
  double ax[NR], ay[NR];
#pragma unroll
  for (int r = 0; r < NR; ++r) { ax[r] = 0.0; ay[r] = 0.0; }
  const bool live = row < n && (!active || active[row]);    // rows of a design that takes no part: nothing read, nothing written
  if (live) {
    const double sigma = sigma_node[row];
    for (int32_t z = rowptr[row] + lane; z < rowptr[row + 1]; z += TPR) {
      const double sm = sigma * vals[(int64_t)S_MINV * nnz + z];
      const double kxx = vals[(int64_t)S_AXX * nnz + z] - sm, kxy = vals[(int64_t)S_AXY * nnz + z];
      const double kyx = vals[(int64_t)S_AYX * nnz + z], kyy = vals[(int64_t)S_AYY * nnz + z] - sm;
      const double* xc = x + 2 * (int64_t)col[z] * NR;
#pragma unroll
      for (int r = 0; r < NR; ++r) {
        const double vx = xc[r], vy = xc[NR + r];
        ax[r] = fma(kxx, vx, ax[r]); ax[r] = fma(kxy, vy, ax[r]);
        ay[r] = fma(kyx, vx, ay[r]); ay[r] = fma(kyy, vy, ay[r]);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < NR; ++r) {
#pragma unroll
    for (int off = TPR / 2; off > 0; off >>= 1) {
      ax[r] += __shfl_down_sync(0xffffffffu, ax[r], off, TPR);
      ay[r] += __shfl_down_sync(0xffffffffu, ay[r], off, TPR);
    }
    if (live && lane == 0) {
      const int64_t o = 2 * row * NR + r;
      t[o] = b[o] - ax[r];
      t[o + NR] = b[o + NR] - ay[r];
    }
  }
}

// out[b][slice] = (sum a^2, sum b^2) over the rows of design b of two interleaved blocks of P right-hand sides
__global__ void __launch_bounds__(256) sq_norms_kernel(const double* __restrict__ a, const double* __restrict__ bvec,
                                                       const int64_t* __restrict__ moff, double* __restrict__ out) {
  __shared__ double sh[64];
  const int b = blockIdx.y, sl = blockIdx.x;
  const int64_t m0 = moff[b] * SOLVE_NRHS, mlen = (moff[b + 1] - moff[b]) * SOLVE_NRHS;
  const int64_t chunk = (mlen + gridDim.x - 1) / gridDim.x;
  const int64_t i0 = m0 + sl * chunk, i1 = min(m0 + mlen, i0 + chunk);
  double sa = 0.0, sb = 0.0;
  for (int64_t i = i0 + threadIdx.x; i < i1; i += 256) { sa = fma(a[i], a[i], sa); sb = fma(bvec[i], bvec[i], sb); }
  sa = block_sum(sa, sh);
  __syncthreads();
  sb = block_sum(sb, sh + 32);
  if (threadIdx.x == 0) { out[(b * gridDim.x + sl) * 2] = sa; out[(b * gridDim.x + sl) * 2 + 1] = sb; }
}

// block of P columns (leading dimension ld) -> interleaved right-hand sides
__global__ void __launch_bounds__(256) interleave_kernel(const double* __restrict__ in, int64_t ld, int64_t m, double* __restrict__ out) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= m) return;
#pragma unroll
  for (int r = 0; r < SOLVE_NRHS; ++r) out[i * SOLVE_NRHS + r] = in[r * ld + i];
}

// out (P columns, leading dimension ld) = interleaved x (+ interleaved dx when given)
__global__ void __launch_bounds__(256) deinterleave_add_kernel(const double* __restrict__ x, const double* __restrict__ dx, int64_t ld,
                                                               int64_t m, double* __restrict__ out, const uint8_t* __restrict__ active) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= m) return;
  const bool corr = dx && (!active || active[i >> 1]);     // the refinement correction exists only for the designs that took part
#pragma unroll
  for (int r = 0; r < SOLVE_NRHS; ++r) out[r * ld + i] = x[i * SOLVE_NRHS + r] + (corr ? dx[i * SOLVE_NRHS + r] : 0.0);
}

__global__ void __launch_bounds__(256) add_kernel(double* __restrict__ x, const double* __restrict__ dx, int64_t m) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < m) x[i] += dx[i];
}

// h[c] = <Q[:,c], r> for c < ncols; one CTA per column
__global__ void __launch_bounds__(RED_T) dots_kernel(const double* __restrict__ Q, int64_t ld, const double* __restrict__ r,
                                                     int64_t m, double* __restrict__ h) {
  __shared__ double sh[32];
  const double* q = Q + (int64_t)blockIdx.x * ld;
  double acc = 0.0;
  for (int64_t i = threadIdx.x; i < m; i += RED_T) acc = fma(q[i], r[i], acc);
  acc = block_sum(acc, sh);
  if (threadIdx.x == 0) h[blockIdx.x] = acc;
}

// r -= V[:, 0..ncols) h
__global__ void __launch_bounds__(256) update_kernel(const double* __restrict__ V, int64_t ld, const double* __restrict__ h,
                                                     int ncols, int64_t m, double* __restrict__ r) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= m) return;
  double acc0 = 0.0, acc1 = 0.0;
  int c = 0;
  for (; c + 1 < ncols; c += 2) {
    acc0 = fma(V[(int64_t)c * ld + i], h[c], acc0);
    acc1 = fma(V[(int64_t)(c + 1) * ld + i], h[c + 1], acc1);
  }
  if (c < ncols) acc0 = fma(V[(int64_t)c * ld + i], h[c], acc0);
  r[i] -= acc0 + acc1;
}

// beta2 = <r, u>; alpha[j] = h1[j] + h2[j]; beta[j] = sqrt(beta2)     (single CTA)
__global__ void __launch_bounds__(RED_T) bnorm_kernel(const double* __restrict__ r, const double* __restrict__ u, int64_t m,
                                                      const double* __restrict__ h1, const double* __restrict__ h2, int j,
                                                      double* __restrict__ alpha, double* __restrict__ beta) {
  __shared__ double sh[32];
  double acc = 0.0;
  for (int64_t i = threadIdx.x; i < m; i += RED_T) acc = fma(r[i], u[i], acc);
  acc = block_sum(acc, sh);
  if (threadIdx.x == 0) {
    beta[j] = sqrt(acc);
    if (h1) alpha[j] = h1[j] + h2[j];
  }
}

// v = r / beta[j], bv = u / beta[j]
__global__ void __launch_bounds__(256) scale_kernel(const double* __restrict__ r, const double* __restrict__ u, int64_t m,
                                                    const double* __restrict__ beta, int j, double* __restrict__ v,
                                                    double* __restrict__ bv) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= m) return;
  const double s = 1.0 / beta[j];
  v[i] = r[i] * s;
  bv[i] = u[i] * s;
}

// Out[:, c] = sum_j In[:, j] * S[j, c]  for c < nout (S is nin x nout, column-major, in global memory)
__global__ void __launch_bounds__(128) rotate_kernel(const double* __restrict__ In, int64_t ld, int nin,
                                                     const double* __restrict__ S, int nout, int64_t m,
                                                     double* __restrict__ Out, int64_t ldo) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= m) return;
  for (int c0 = 0; c0 < nout; c0 += 8) {
    double acc[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] = 0.0;
    for (int j = 0; j < nin; ++j) {
      const double v = In[(int64_t)j * ld + i];
#pragma unroll
      for (int c = 0; c < 8; ++c)
        if (c0 + c < nout) acc[c] = fma(v, __ldg(S + (int64_t)(c0 + c) * nin + j), acc[c]);
    }
#pragma unroll
    for (int c = 0; c < 8; ++c)
      if (c0 + c < nout) Out[(int64_t)(c0 + c) * ldo + i] = acc[c];
  }
}

// Out_b[:, c] = sum_j In_b[:, j] * S_b[j, c]  for c < nout, every design of the forest in one launch (S_b at S + b*sstride).
// A thread owns a PAIR of rows (128-bit loads) and 16 output columns at a time; the 16 columns of S sit in shared memory as
// [j][16], read as broadcast 128-bit loads: 32 FMAs per 9 loads (the first version: 8 FMAs per 9 loads, In re-read per 8 columns).
constexpr int ROT_C = 16;        // output columns per pass
constexpr int ROT_MAXIN = 256;   // basis columns the shared tile of S holds (ncv + P <= 3 (n_modes + 12) + 8)
__global__ void __launch_bounds__(128) rotate_forest_kernel(const double* __restrict__ In, int64_t ld, int nin,
                                                            const double* __restrict__ S, int64_t sstride, int nout,
                                                            const int64_t* __restrict__ moff, double* __restrict__ Out, int64_t ldo) {
  __shared__ __align__(16) double sh[ROT_MAXIN * ROT_C];
  const int b = blockIdx.y;
  const int64_t i = moff[b] + 2 * (blockIdx.x * (int64_t)blockDim.x + threadIdx.x);
  const bool live = i < moff[b + 1];
  const double* Sb = S + b * sstride;
  for (int c0 = 0; c0 < nout; c0 += ROT_C) {
    __syncthreads();
    for (int t = threadIdx.x; t < nin * ROT_C; t += blockDim.x) {
      const int c = t / nin, j = t - c * nin;
      sh[j * ROT_C + c] = c0 + c < nout ? Sb[(int64_t)(c0 + c) * nin + j] : 0.0;
    }
    __syncthreads();
    if (!live) continue;
    double2 acc[ROT_C];
#pragma unroll
    for (int c = 0; c < ROT_C; ++c) acc[c] = make_double2(0.0, 0.0);
#pragma unroll 2
    for (int j = 0; j < nin; ++j) {
      const double2 v = *reinterpret_cast<const double2*>(In + (int64_t)j * ld + i);
      const double2* sj = reinterpret_cast<const double2*>(sh + j * ROT_C);
#pragma unroll
      for (int c = 0; c < ROT_C; c += 2) {
        const double2 w = sj[c >> 1];
        acc[c].x = fma(v.x, w.x, acc[c].x); acc[c].y = fma(v.y, w.x, acc[c].y);
        acc[c + 1].x = fma(v.x, w.y, acc[c + 1].x); acc[c + 1].y = fma(v.y, w.y, acc[c + 1].y);
      }
    }
#pragma unroll
    for (int c = 0; c < ROT_C; ++c)
      if (c0 + c < nout) *reinterpret_cast<double2*>(Out + (int64_t)(c0 + c) * ldo + i) = acc[c];
  }
}

// ---- block Lanczos pieces (P = SOLVE_NRHS vectors per block), batched over the designs of a forest ------------
// All vectors are concatenations over the designs: design b owns rows [moff[b], moff[b+1]).  Everything that
// couples rows — inner products, the small Cholesky, the coefficient matrices — is per design (blockIdx.y);
// with one design the kernels do exactly what the single-design versions did.
constexpr int P = SOLVE_NRHS;

// Deterministic CTA reduction of N accumulators at once: shuffles inside the warps, one trip through shared memory,
// one barrier.  sh holds N * (blockDim.x / 32) doubles; thread t < N returns the total of accumulator t.
template <int N>
__device__ __forceinline__ double block_sum_n(double (&acc)[N], double* sh) {
  const int lane = threadIdx.x % 32, warp = threadIdx.x / 32, nw = blockDim.x / 32;
#pragma unroll
  for (int a = 0; a < N; ++a) {
    double v = acc[a];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
    if (lane == 0) sh[a * nw + warp] = v;
  }
  __syncthreads();
  double t = 0.0;
  if ((int)threadIdx.x < N)
    for (int w = 0; w < nw; ++w) t += sh[threadIdx.x * nw + w];
  return t;
}

constexpr int DCOLS = 8;    // basis columns per CTA of the projection kernel: the block R is read once per 8 columns
constexpr int RSPLIT = 8;   // row slices per design (fixed: the summation order does not depend on the forest)

// Hp_b[s][c + r*ldh] = <Q_b[rows of slice s, c], R_b[rows of slice s, r]> for c < ncols, r < P;
// one CTA per (group of DCOLS columns, row slice s, design b).  Rows are taken in PAIRS (128-bit loads: the design's rows
// start at an even offset and the slices are cut at even rows), two pairs in flight per thread.
__global__ void __launch_bounds__(256) dots_block_kernel(const double* __restrict__ Q, int64_t ld, const double* __restrict__ R,
                                                         const int64_t* __restrict__ moff, int ncols, double* __restrict__ Hp, int ldh,
                                                         int64_t hstride) {
  __shared__ double sh[DCOLS * P * 8];
  const int b = blockIdx.z, sl = blockIdx.y, c0 = blockIdx.x * DCOLS;
  const int64_t m0 = moff[b], mlen = moff[b + 1] - m0;
  const int64_t chunk = 2 * ((mlen / 2 + RSPLIT - 1) / RSPLIT);
  const int64_t i0 = m0 + sl * chunk, i1 = min(m0 + mlen, i0 + chunk);
  const int nc = min(DCOLS, ncols - c0);
  const double* q = Q + (int64_t)c0 * ld;
  double acc[DCOLS * P];
#pragma unroll
  for (int a = 0; a < DCOLS * P; ++a) acc[a] = 0.0;
  for (int64_t i = i0 + 2 * threadIdx.x; i < i1; i += 512) {
    double2 rv[P], qv[DCOLS];
#pragma unroll
    for (int r = 0; r < P; ++r) rv[r] = *reinterpret_cast<const double2*>(R + r * ld + i);
#pragma unroll
    for (int c = 0; c < DCOLS; ++c) qv[c] = c < nc ? *reinterpret_cast<const double2*>(q + c * ld + i) : make_double2(0.0, 0.0);
#pragma unroll
    for (int c = 0; c < DCOLS; ++c)
#pragma unroll
      for (int r = 0; r < P; ++r) acc[c * P + r] = fma(qv[c].y, rv[r].y, fma(qv[c].x, rv[r].x, acc[c * P + r]));
  }
  const double t = block_sum_n<DCOLS * P>(acc, sh);
  const int c = threadIdx.x / P, r = threadIdx.x % P;
  if ((int)threadIdx.x < DCOLS * P && c < nc) Hp[(b * RSPLIT + sl) * hstride + c0 + c + r * ldh] = t;
}

// H_b = sum over the row slices of Hp_b (fixed order)
__global__ void sum_slices_kernel(const double* __restrict__ Hp, int ldh, int64_t hstride, int ncols, double* __restrict__ H) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (c >= ncols) return;
#pragma unroll
  for (int r = 0; r < P; ++r) {
    double t = 0.0;
#pragma unroll
    for (int sl = 0; sl < RSPLIT; ++sl) t += Hp[(b * RSPLIT + sl) * hstride + c + r * ldh];
    H[b * hstride + c + r * ldh] = t;
  }
}

// R_b[:, r] -= V_b[:, 0..ncols) H_b[:, r]; a thread owns a PAIR of rows (128-bit loads), the coefficients sit in shared memory
constexpr int UPD_MAXC = 384;      // basis columns the shared copy of H holds (the basis has ncv + P <= 3 (n_modes + 12) + 8)
__global__ void __launch_bounds__(256) update_block_kernel(const double* __restrict__ V, int64_t ld, const double* __restrict__ H,
                                                           int ldh, int64_t hstride, int ncols, const int64_t* __restrict__ moff,
                                                           double* __restrict__ R) {
  __shared__ double sh[UPD_MAXC * P];     // [c][r]
  const int b = blockIdx.y;
  const double* Hb = H + b * hstride;
  const int nsh = min(ncols, UPD_MAXC);
  for (int t = threadIdx.x; t < nsh * P; t += 256) sh[t] = Hb[(t / P) + (t % P) * ldh];
  __syncthreads();
  const int64_t i = moff[b] + 2 * (blockIdx.x * (int64_t)blockDim.x + threadIdx.x);
  if (i >= moff[b + 1]) return;
  double2 acc[P];
#pragma unroll
  for (int r = 0; r < P; ++r) acc[r] = make_double2(0.0, 0.0);
  int c = 0;
#pragma unroll 8
  for (; c < nsh; ++c) {
    const double2 v = *reinterpret_cast<const double2*>(V + (int64_t)c * ld + i);
#pragma unroll
    for (int r = 0; r < P; ++r) { const double h = sh[c * P + r]; acc[r].x = fma(v.x, h, acc[r].x); acc[r].y = fma(v.y, h, acc[r].y); }
  }
  for (; c < ncols; ++c) {           // beyond the shared copy (never with the basis sizes of this solver)
    const double2 v = *reinterpret_cast<const double2*>(V + (int64_t)c * ld + i);
#pragma unroll
    for (int r = 0; r < P; ++r) { const double h = __ldg(Hb + c + r * ldh); acc[r].x = fma(v.x, h, acc[r].x); acc[r].y = fma(v.y, h, acc[r].y); }
  }
#pragma unroll
  for (int r = 0; r < P; ++r) {
    double2* p = reinterpret_cast<double2*>(R + r * ld + i);
    double2 o = *p;
    o.x -= acc[r].x; o.y -= acc[r].y;
    *p = o;
  }
}

// Hs_b[0..ncols, r] = h1 + h2 (the projected-matrix column block kept for the host)
// (the first pass covered the columns c >= c1 only: h1 counts from there)
__global__ void store_h_kernel(const double* __restrict__ h1, const double* __restrict__ h2, int ldh, int64_t hstride, int ncols, int c1,
                               double* __restrict__ Hs, int lds, int64_t sstride) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (c >= ncols) return;
#pragma unroll
  for (int r = 0; r < P; ++r)
    Hs[b * sstride + c + r * lds] = (c >= c1 ? h1[b * hstride + c + r * ldh] : 0.0) + h2[b * hstride + c + r * ldh];
}

// Gp_b[s][r + q*P] = <R_b[slice s, r], U_b[slice s, q]>, one CTA per (row slice, design): R and U are read once
__global__ void __launch_bounds__(256) gram_kernel(const double* __restrict__ R, const double* __restrict__ U, int64_t ld,
                                                   const int64_t* __restrict__ moff, double* __restrict__ Gp) {
  __shared__ double sh[P * P * 8];
  const int b = blockIdx.y, sl = blockIdx.x;
  const int64_t m0 = moff[b], mlen = moff[b + 1] - m0;
  const int64_t chunk = (mlen + RSPLIT - 1) / RSPLIT;
  const int64_t i0 = m0 + sl * chunk, i1 = min(m0 + mlen, i0 + chunk);
  double acc[P * P];
#pragma unroll
  for (int a = 0; a < P * P; ++a) acc[a] = 0.0;
  for (int64_t i = i0 + threadIdx.x; i < i1; i += 256) {
    double rv[P], uv[P];
#pragma unroll
    for (int r = 0; r < P; ++r) { rv[r] = R[r * ld + i]; uv[r] = U[r * ld + i]; }
#pragma unroll
    for (int q = 0; q < P; ++q)
#pragma unroll
      for (int r = 0; r < P; ++r) acc[r + q * P] = fma(rv[r], uv[q], acc[r + q * P]);
  }
  const double t = block_sum_n<P * P>(acc, sh);
  if ((int)threadIdx.x < P * P) Gp[(b * RSPLIT + sl) * P * P + threadIdx.x] = t;
}

// Cholesky G_b = L L^T (P x P, G_b = sum of the slices of Gp_b, symmetrised), Linv = L^-1, one CTA per design.  Lout (for
// the host): L, column-major, slot `slot` of design b.  cstat[b] = 1 on breakdown (the residual block of that design
// lost rank).
__global__ void chol_kernel(const double* __restrict__ Gp, double* __restrict__ Lall, int nslots, int slot, double* __restrict__ Linv,
                            int32_t* __restrict__ cstat) {
  if (threadIdx.x != 0) return;
  const int b = blockIdx.x;
  Linv += b * P * P;
  double* Lout = Lall + ((size_t)b * nslots + slot) * P * P;
  double G[P * P];
  for (int a = 0; a < P * P; ++a) {
    double t = 0.0;
    for (int sl = 0; sl < RSPLIT; ++sl) t += Gp[(b * RSPLIT + sl) * P * P + a];
    G[a] = t;
  }
  double A[P][P], L[P][P], Li[P][P];
  for (int i = 0; i < P; ++i) for (int j = 0; j < P; ++j) { A[i][j] = 0.5 * (G[i + j * P] + G[j + i * P]); L[i][j] = 0.0; Li[i][j] = 0.0; }
  double dmax = 0.0;
  for (int i = 0; i < P; ++i) dmax = fmax(dmax, A[i][i]);
  for (int j = 0; j < P; ++j) {
    double d = A[j][j];
    for (int k = 0; k < j; ++k) d -= L[j][k] * L[j][k];
    if (!(d > 1e-24 * dmax) || !isfinite(d)) { cstat[b] = 1; d = 1e-24 * dmax + 1e-300; }
    L[j][j] = sqrt(d);
    for (int i = j + 1; i < P; ++i) {
      double v = A[i][j];
      for (int k = 0; k < j; ++k) v -= L[i][k] * L[j][k];
      L[i][j] = v / L[j][j];
    }
  }
  for (int j = 0; j < P; ++j) {           // Li = L^-1 by forward substitution on the identity
    for (int i = 0; i < P; ++i) {
      double v = (i == j) ? 1.0 : 0.0;
      for (int k = 0; k < i; ++k) v -= L[i][k] * Li[k][j];
      Li[i][j] = v / L[i][i];
    }
  }
  for (int i = 0; i < P; ++i) for (int j = 0; j < P; ++j) { Lout[i + j * P] = L[i][j]; Linv[i + j * P] = Li[i][j]; }
}

// Vn_b = R_b L_b^-T, BVn_b = U_b L_b^-T   (Vn[:, r] = sum_s R[:, s] * Linv[r, s])
__global__ void __launch_bounds__(256) scale_block_kernel(const double* __restrict__ R, const double* __restrict__ U, int64_t ld,
                                                          const int64_t* __restrict__ moff, const double* __restrict__ Linv,
                                                          double* __restrict__ Vn, double* __restrict__ BVn) {
  const int b = blockIdx.y;
  const int64_t i = moff[b] + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= moff[b + 1]) return;
  const double* Lb = Linv + b * P * P;
  double rv[P], uv[P];
#pragma unroll
  for (int s2 = 0; s2 < P; ++s2) { rv[s2] = R[s2 * ld + i]; uv[s2] = U[s2 * ld + i]; }
#pragma unroll
  for (int r = 0; r < P; ++r) {
    double a = 0.0, c = 0.0;
#pragma unroll
    for (int s2 = 0; s2 <= r; ++s2) { const double l = __ldg(Lb + r + s2 * P); a = fma(rv[s2], l, a); c = fma(uv[s2], l, c); }
    Vn[r * ld + i] = a; BVn[r * ld + i] = c;
  }
}

// deterministic pseudo-random start vectors for block columns 1..P-1 (column 0 is the caller's v0 / ones); the
// sequence depends on the row index WITHIN the design, so a design starts the same alone or in a forest
__global__ void start_block_kernel(double* __restrict__ R, int64_t ld, const int64_t* __restrict__ moff) {
  const int b = blockIdx.y;
  const int64_t il = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t i = moff[b] + il;
  if (i >= moff[b + 1]) return;
  for (int r = 1; r < P; ++r) {
    uint64_t h = (uint64_t)il * 0x9E3779B97F4A7C15ull + (uint64_t)r * 0xBF58476D1CE4E5B9ull + 0x94D049BB133111EBull;
    h ^= h >> 30; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 27; h *= 0x94D049BB133111EBull; h ^= h >> 31;
    R[r * ld + i] = (double)(h >> 11) * (2.0 / 9007199254740992.0) - 1.0;
  }
}

__global__ void fill_kernel(double* v, int64_t m, double val) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < m) v[i] = val;
}

// ---- per-mode reductions ----------------------------------------------------------------------------------
constexpr int NRED = 10;  // norm2, e_core, px_core, py_core, px_all, py_all, div, res2, ||B||_F^2, ||A||_F^2
constexpr int MROWS = 2048;  // rows (nodes) per CTA in the mode reduction

// Four threads per row (consecutive non-zeros across the lanes: coalesced value loads, no divergence between vertex rows of 19
// and edge rows of 9 entries) and TWO modes per pass over the matrix (the eight value arrays are read once for both; the first
// version read them once per mode with one thread per row: 154 us per 7-core design).
constexpr int MODE_TPR = 4, MODE_PAIR = 2;
__global__ void __launch_bounds__(256) mode_partial_kernel(int32_t row0, int32_t row1, const int32_t* __restrict__ rowptr,
                                                           const int32_t* __restrict__ col, const double* __restrict__ vals,
                                                           int64_t nnz, const uint8_t* __restrict__ in_core,
                                                           const double2* __restrict__ X, int64_t ldx /* in double2 */,
                                                           const double* __restrict__ lambda, int nmodes, double* __restrict__ part,
                                                           int nchunks) {
  __shared__ double sh[32];
  const int mode0 = blockIdx.y * MODE_PAIR, chunk = blockIdx.x;
  const int sub = threadIdx.x % MODE_TPR;
  double acc[MODE_PAIR][NRED];
#pragma unroll
  for (int q = 0; q < MODE_PAIR; ++q)
#pragma unroll
    for (int k = 0; k < NRED; ++k) acc[q][k] = 0.0;
  const double2* x[MODE_PAIR];
  double lam[MODE_PAIR];
#pragma unroll
  for (int q = 0; q < MODE_PAIR; ++q) {
    const int md = min(mode0 + q, nmodes - 1);          // an odd count: the last pass computes its mode twice, stores it once
    x[q] = X + (int64_t)md * ldx; lam[q] = lambda[md];
  }
  const int r0 = row0 + chunk * MROWS, r1 = min(row1, r0 + MROWS);   // rows of ONE design of the forest
  for (int rb = r0 + (int)threadIdx.x / MODE_TPR; rb - (int)threadIdx.x / MODE_TPR < r1; rb += 256 / MODE_TPR) {
    const int r = rb;
    const bool live = r < r1;                             // whole groups of four stay in the shuffles below
    double s[MODE_PAIR][6];                               // dx dy ax ay bx by of this lane's share of the row
#pragma unroll
    for (int q = 0; q < MODE_PAIR; ++q)
#pragma unroll
      for (int k = 0; k < 6; ++k) s[q][k] = 0.0;
    double fa = 0.0, fb = 0.0;
    if (live) {
      for (int32_t z = rowptr[r] + sub; z < rowptr[r + 1]; z += MODE_TPR) {
        const int32_t cz = col[z];
        const double dxx = vals[(int64_t)S_DXX * nnz + z], dxy = vals[(int64_t)S_DXY * nnz + z], dyy = vals[(int64_t)S_DYY * nnz + z];
        const double a0 = vals[(int64_t)S_AXX * nnz + z], a1 = vals[(int64_t)S_AXY * nnz + z];
        const double a2 = vals[(int64_t)S_AYX * nnz + z], a3 = vals[(int64_t)S_AYY * nnz + z];
        const double mi = vals[(int64_t)S_MINV * nnz + z];
        // squared Frobenius norms of A and B for the normwise backward error
        fa += a0 * a0 + a1 * a1 + a2 * a2 + a3 * a3;
        fb += 2.0 * mi * mi;
#pragma unroll
        for (int q = 0; q < MODE_PAIR; ++q) {
          const double2 c = x[q][cz];
          s[q][0] = fma(dxx, c.x, s[q][0]);
          s[q][0] = fma(2.0 * dxy, c.y, s[q][0]);
          s[q][1] = fma(dyy, c.y, s[q][1]);
          s[q][2] = fma(a0, c.x, s[q][2]);
          s[q][2] = fma(a1, c.y, s[q][2]);
          s[q][3] = fma(a2, c.x, s[q][3]);
          s[q][3] = fma(a3, c.y, s[q][3]);
          s[q][4] = fma(mi, c.x, s[q][4]); s[q][5] = fma(mi, c.y, s[q][5]);
        }
      }
    }
    // the row's sums (the residual is squared per ROW): fixed butterfly over the four lanes
#pragma unroll
    for (int q = 0; q < MODE_PAIR; ++q)
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        s[q][k] += __shfl_xor_sync(0xffffffffu, s[q][k], 1);
        s[q][k] += __shfl_xor_sync(0xffffffffu, s[q][k], 2);
      }
    if (live && sub == 0) {
      const bool core = in_core[r];
#pragma unroll
      for (int q = 0; q < MODE_PAIR; ++q) {
        const double2 v = x[q][r];
        const double ex = v.x * v.x, ey = v.y * v.y;
        acc[q][0] += ex + ey;
        acc[q][4] += ex; acc[q][5] += ey;
        if (core) { acc[q][1] += ex + ey; acc[q][2] += ex; acc[q][3] += ey; }
        acc[q][6] += v.x * s[q][0] + v.y * s[q][1];
        const double rx = s[q][2] - lam[q] * s[q][4], ry = s[q][3] - lam[q] * s[q][5];
        acc[q][7] += rx * rx + ry * ry;
      }
    }
    if (live) {                                           // every lane's share of the Frobenius norms
#pragma unroll
      for (int q = 0; q < MODE_PAIR; ++q) { acc[q][8] += fb; acc[q][9] += fa; }
    }
  }
#pragma unroll
  for (int q = 0; q < MODE_PAIR; ++q) {
    if (mode0 + q >= nmodes) break;
    for (int k = 0; k < NRED; ++k) {
      const double t = block_sum(acc[q][k], sh);
      if (threadIdx.x == 0) part[((int64_t)(mode0 + q) * nchunks + chunk) * NRED + k] = t;
    }
  }
}

// metrics (k, 8): div_energy, sum_e_core, sum_e, Px_core, Py_core, Px_all, Py_all, norm2_raw ; resid (k,2)
__global__ void mode_final_kernel(const double* __restrict__ part, int nchunks, int k, const double* __restrict__ lambda,
                                  double* __restrict__ metrics, double* __restrict__ resid, double* __restrict__ scale) {
  const int mode = blockIdx.x * blockDim.x + threadIdx.x;
  if (mode >= k) return;
  double t[NRED];
  for (int q = 0; q < NRED; ++q) t[q] = 0.0;
  for (int c = 0; c < nchunks; ++c)
    for (int q = 0; q < NRED; ++q) t[q] += part[((int64_t)mode * nchunks + c) * NRED + q];
  const double nrm = sqrt(t[0]) + 1e-30;
  const double inv2 = 1.0 / (nrm * nrm);
  scale[mode] = 1.0 / nrm;
  double* o = metrics + (int64_t)mode * PLFEM_NMETRICS;
  o[0] = t[6] * inv2; o[1] = t[1] * inv2; o[2] = t[0] * inv2; o[3] = t[2] * inv2; o[4] = t[3] * inv2;
  o[5] = t[4] * inv2; o[6] = t[5] * inv2; o[7] = t[0];
  // ||A x - lambda B x||_2 and (||A||_F + |lambda| ||B||_F) ||x||_2
  resid[2 * mode] = sqrt(t[7]); resid[2 * mode + 1] = (sqrt(t[9]) + fabs(lambda[mode]) * sqrt(t[8])) * sqrt(t[0]);
}

__global__ void __launch_bounds__(256) write_evecs_kernel(int32_t row0, int32_t n, const int32_t* __restrict__ perm,
                                                          const double2* __restrict__ X, int64_t ldx,
                                                          const double* __restrict__ scale, double* __restrict__ out) {
  const int mode = blockIdx.y;
  const int32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const double2 v = X[(int64_t)mode * ldx + row0 + r];
  const double s = scale[mode];
  double* o = out + (int64_t)mode * 2 * n;
  const int32_t ip = perm[row0 + r];
  o[ip] = v.x * s;
  o[n + ip] = v.y * s;
}

}  // namespace

void launch_spmm_b(plfem_ctx* ctx, const DevPattern& pat, const double* d_vals, const double* x, double* y, int nrhs, int64_t ld) {
  const unsigned g = (unsigned)(((int64_t)pat.n * 4 + 255) / 256);
  const double* minv = d_vals + (int64_t)S_MINV * pat.nnz;
  if (nrhs == 1) spmm_b_kernel<1><<<g, 256, 0, ctx->stream>>>(pat.n, pat.rowptr.p, pat.col.p, minv, (const double2*)x, (double2*)y, 0);
  else spmm_b_kernel<SOLVE_NRHS><<<g, 256, 0, ctx->stream>>>(pat.n, pat.rowptr.p, pat.col.p, minv, (const double2*)x, (double2*)y, ld / 2);
  PLFEM_CUDA(cudaGetLastError());
  ctx->launches++;
}

void launch_resid_k(plfem_ctx* ctx, const DevPattern& pat, const double* d_vals, const double* d_sigma_node, const double* x,
                    const double* b, double* t, int nrhs, const uint8_t* active) {
  const unsigned g = (unsigned)(((int64_t)pat.n * 4 + 255) / 256);
  if (nrhs == 1)
    resid_k_kernel<1><<<g, 256, 0, ctx->stream>>>(pat.n, pat.rowptr.p, pat.col.p, d_vals, pat.nnz, d_sigma_node, x, b, t, active);
  else
    resid_k_kernel<SOLVE_NRHS><<<g, 256, 0, ctx->stream>>>(pat.n, pat.rowptr.p, pat.col.p, d_vals, pat.nnz, d_sigma_node, x, b, t, active);
  PLFEM_CUDA(cudaGetLastError());
  ctx->launches++;
}

void launch_axpy(plfem_ctx* ctx, double* x, const double* dx, int64_t m) {
  add_kernel<<<(unsigned)((m + 255) / 256), 256, 0, ctx->stream>>>(x, dx, m);
  PLFEM_CUDA(cudaGetLastError());
  ctx->launches++;
}

void run_eigensolver(plfem_ctx* ctx, const DevPattern& pat, DevPlan& D, const double* d_vals, const double* d_sigma_node, double sigma,
                     int k, int ncv, double tol, int maxiter, int refine_steps, const double* d_v0, DevBuf<double>& X,
                     std::vector<double>& lambda, EigenResult& res) {
  const int64_t m = 2 * (int64_t)pat.n;
  const int64_t ld = m;
  cudaStream_t st = ctx->stream;
  DevBuf<double> V[2], BV[2], r, u, h1, h2, alpha, beta, Sdev, rt, rdx;
  for (int b = 0; b < 2; ++b) { V[b].alloc(ctx, (size_t)ld * (ncv + 1)); BV[b].alloc(ctx, (size_t)ld * (ncv + 1)); }
  r.alloc(ctx, m); u.alloc(ctx, m);
  rt.alloc(ctx, m); rdx.alloc(ctx, m);
  h1.alloc(ctx, ncv + 1); h2.alloc(ctx, ncv + 1); alpha.alloc(ctx, ncv + 1); beta.alloc(ctx, ncv + 1);
  Sdev.alloc(ctx, (size_t)ncv * ncv);
  int cur = 0;
  const unsigned gm = (unsigned)((m + 255) / 256);
  auto spmm = [&](const double* x, double* y) {
    launch_spmm_b(ctx, pat, d_vals, x, y);
  };

  // OP application: block-LDL^T solve + fixed number of refinement steps with the true operator
  // (the pivot blocks are not pivoted against each other, so one raw solve carries a backward error
  // of ~1e-9 relative; each step squares the contraction and the composite stays a fixed symmetric
  // linear operator, which is what Lanczos needs).  The whole application — 2 sweeps x levels x
  // (1 + refine) launches — is captured once into a CUDA graph on fixed buffers (opin -> r) and
  // replayed, so the CPU issues one launch per operator application instead of ~70.
  const bool use_graph = true;
  DevBuf<double> opin;
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t gexec = nullptr;
  int graph_nodes = 0;
  if (use_graph) {
    opin.alloc(ctx, m);
    const int before = ctx->launches;
    PLFEM_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    try {
      run_solve(ctx, D, opin.p, r.p);
      for (int it = 0; it < refine_steps; ++it) {
        launch_resid_k(ctx, pat, d_vals, d_sigma_node, r.p, opin.p, rt.p);
        run_solve(ctx, D, rt.p, rdx.p);
        launch_axpy(ctx, r.p, rdx.p, m);
      }
    } catch (...) {
      cudaGraph_t dead = nullptr;
      cudaStreamEndCapture(st, &dead);
      if (dead) cudaGraphDestroy(dead);
      throw;
    }
    PLFEM_CUDA(cudaStreamEndCapture(st, &graph));
    graph_nodes = ctx->launches - before;
    ctx->launches = before;
    PLFEM_CUDA(cudaGraphInstantiate(&gexec, graph, 0));
  }
  struct GraphGuard {
    cudaGraph_t g; cudaGraphExec_t e;
    ~GraphGuard() { if (e) cudaGraphExecDestroy(e); if (g) cudaGraphDestroy(g); }
  } guard{graph, gexec};
  auto apply_op = [&](const double* bvec) {   // r = OP(bvec)
    if (use_graph) {
      PLFEM_CUDA(cudaMemcpyAsync(opin.p, bvec, m * sizeof(double), cudaMemcpyDeviceToDevice, st));
      PLFEM_CUDA(cudaGraphLaunch(gexec, st));
      ctx->launches += graph_nodes;
    }
  };

  // start vector: v0 / ||v0||_B
  if (d_v0) PLFEM_CUDA(cudaMemcpyAsync(r.p, d_v0, m * sizeof(double), cudaMemcpyDeviceToDevice, st));
  else { fill_kernel<<<gm, 256, 0, st>>>(r.p, m, 1.0); ctx->launches++; }
  spmm(r.p, u.p);
  bnorm_kernel<<<1, RED_T, 0, st>>>(r.p, u.p, m, nullptr, nullptr, ncv, alpha.p, beta.p);   // beta[ncv] = ||v0||_B
  scale_kernel<<<gm, 256, 0, st>>>(r.p, u.p, m, beta.p, ncv, V[cur].p, BV[cur].p);
  ctx->launches += 2;

  std::vector<double> h_alpha(ncv + 1), h_beta(ncv + 1), theta_keep, b_keep, T, w;
  std::vector<int> order(ncv);
  int p = 0;
  res = EigenResult();
  const double eps23 = std::pow(2.220446049250313e-16, 2.0 / 3.0);
  for (;;) {
    for (int j = p; j < ncv; ++j) {
      double* Vc = V[cur].p; double* BVc = BV[cur].p;
      apply_op(BVc + (int64_t)j * ld);                                        // r = OP v_j
      res.n_op++;
      dots_kernel<<<j + 1, RED_T, 0, st>>>(BVc, ld, r.p, m, h1.p);            // CGS pass 1
      update_kernel<<<gm, 256, 0, st>>>(Vc, ld, h1.p, j + 1, m, r.p);
      dots_kernel<<<j + 1, RED_T, 0, st>>>(BVc, ld, r.p, m, h2.p);            // CGS pass 2
      update_kernel<<<gm, 256, 0, st>>>(Vc, ld, h2.p, j + 1, m, r.p);
      spmm(r.p, u.p);
      bnorm_kernel<<<1, RED_T, 0, st>>>(r.p, u.p, m, h1.p, h2.p, j, alpha.p, beta.p);
      scale_kernel<<<gm, 256, 0, st>>>(r.p, u.p, m, beta.p, j, Vc + (int64_t)(j + 1) * ld, BVc + (int64_t)(j + 1) * ld);
      ctx->launches += 6;
    }
    PLFEM_CUDA(cudaGetLastError());
    alpha.download(h_alpha.data(), ncv);
    beta.download(h_beta.data(), ncv);
    PLFEM_CUDA(stream_wait(st));
    // projected matrix
    T.assign((size_t)ncv * ncv, 0.0);
    for (int i = 0; i < p; ++i) { T[(size_t)i * ncv + i] = theta_keep[i]; T[(size_t)p * ncv + i] = T[(size_t)i * ncv + p] = b_keep[i]; }
    for (int j = p; j < ncv; ++j) {
      T[(size_t)j * ncv + j] = h_alpha[j];
      if (j + 1 < ncv) T[(size_t)(j + 1) * ncv + j] = T[(size_t)j * ncv + j + 1] = h_beta[j];
    }
    for (double v : T) if (!std::isfinite(v)) throw StatusError(PLFEM_ERR_SINGULAR, "Lanczos recurrence produced a non-finite value (shifted operator singular?)");
    symmetric_eigen(ncv, T, w);   // T now holds eigenvectors in columns
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return std::fabs(w[a]) > std::fabs(w[b]); });
    const double bm = h_beta[ncv - 1];
    int nconv = 0;
    for (int i = 0; i < k; ++i) {
      const int c = order[i];
      const double bound = std::fabs(bm * T[(size_t)c * ncv + (ncv - 1)]);
      if (bound <= tol * std::max(eps23, std::fabs(w[c]))) nconv++;
    }
    res.nconv = nconv;
    const bool done = nconv >= k;
    if (done || res.n_restart >= maxiter) {
      // eigenvectors X = V S[:, wanted], eigenvalues lambda = sigma + 1/theta, ascending in lambda
      std::vector<int> sel(order.begin(), order.begin() + k);
      std::sort(sel.begin(), sel.end(), [&](int a, int b) { return sigma + 1.0 / w[a] < sigma + 1.0 / w[b]; });
      std::vector<double> S((size_t)ncv * k);
      lambda.resize(k); res.theta.resize(k);
      for (int i = 0; i < k; ++i) {
        std::copy(T.begin() + (size_t)sel[i] * ncv, T.begin() + (size_t)(sel[i] + 1) * ncv, S.begin() + (size_t)i * ncv);
        res.theta[i] = w[sel[i]];
        lambda[i] = sigma + 1.0 / w[sel[i]];
      }
      PLFEM_CUDA(cudaMemcpyAsync(Sdev.p, S.data(), S.size() * sizeof(double), cudaMemcpyHostToDevice, st));
      X.alloc(ctx, (size_t)m * k);
      rotate_kernel<<<(unsigned)((m + 127) / 128), 128, 0, st>>>(V[cur].p, ld, ncv, Sdev.p, k, m, X.p, m);
      ctx->launches++;
      PLFEM_CUDA(stream_wait(st));   // S is a local host buffer
      if (!done) throw StatusError(PLFEM_ERR_NO_CONVERGENCE, "Lanczos: " + std::to_string(nconv) + " of " + std::to_string(k) + " eigenpairs converged after " + std::to_string(res.n_restart) + " restarts");
      return;
    }
    // thick restart: keep the k wanted Ritz pairs plus some of the next ones (ARPACK's nev adjustment)
    int keep = k + std::min(nconv, (ncv - k) / 2);
    keep = std::max(1, std::min(keep, ncv - 2));
    std::vector<double> S((size_t)ncv * keep);
    theta_keep.resize(keep); b_keep.resize(keep);
    for (int i = 0; i < keep; ++i) {
      const int c = order[i];
      std::copy(T.begin() + (size_t)c * ncv, T.begin() + (size_t)(c + 1) * ncv, S.begin() + (size_t)i * ncv);
      theta_keep[i] = w[c];
      b_keep[i] = bm * T[(size_t)c * ncv + (ncv - 1)];
    }
    PLFEM_CUDA(cudaMemcpyAsync(Sdev.p, S.data(), S.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    const int nxt = cur ^ 1;
    rotate_kernel<<<(unsigned)((m + 127) / 128), 128, 0, st>>>(V[cur].p, ld, ncv, Sdev.p, keep, m, V[nxt].p, ld);
    rotate_kernel<<<(unsigned)((m + 127) / 128), 128, 0, st>>>(BV[cur].p, ld, ncv, Sdev.p, keep, m, BV[nxt].p, ld);
    PLFEM_CUDA(cudaMemcpyAsync(V[nxt].p + (int64_t)keep * ld, V[cur].p + (int64_t)ncv * ld, m * sizeof(double), cudaMemcpyDeviceToDevice, st));
    PLFEM_CUDA(cudaMemcpyAsync(BV[nxt].p + (int64_t)keep * ld, BV[cur].p + (int64_t)ncv * ld, m * sizeof(double), cudaMemcpyDeviceToDevice, st));
    ctx->launches += 2;
    PLFEM_CUDA(stream_wait(st));     // S is a local host buffer
    cur = nxt; p = keep;
    res.n_restart++;
  }
}

// ---- thick-restart BLOCK Lanczos (block size P) on a forest of designs ---------------------------------------
// Same operator, same inner product, same convergence test as the single-vector version, but every
// operator application carries P vectors through the sweeps: the factor is read once for P right-hand sides
// and the number of SEQUENTIAL operator applications — what bounds the latency of this phase, each being a
// chain of ~70 dependent small kernels — drops from ~57 to ~24 for config 1.  The projected matrix is built
// from the full-reorthogonalisation coefficients (its upper triangle is exactly what the CGS passes
// produce), so a thick restart needs no special-casing: the couplings between kept Ritz vectors and the
// residual block reappear as the first coefficients computed after the restart.
//
// Forest: the designs of a batch are independent eigenproblems on disjoint row ranges of the same vectors.
// They advance in lockstep — ONE chain of launches serves all of them, which is what turns a latency-bound
// solve into a throughput-bound one — each with its own inner products, projected matrix, convergence test
// and restart rotation.  A design that converges has its Ritz vectors extracted at once; it keeps being
// carried along (its rows cannot influence any other design) until the last one is done.
namespace {
template <class F>
void for_each_design(int nb, F&& fn) {
  const int nt = std::min<int>(nb, std::max(1u, std::thread::hardware_concurrency()));
  if (nt <= 1) { for (int b = 0; b < nb; ++b) fn(b); return; }
  std::atomic<int> next{0};
  std::exception_ptr err; std::mutex mu;
  auto work = [&] {
    for (int b; (b = next.fetch_add(1)) < nb;) {
      try { fn(b); } catch (...) { std::lock_guard<std::mutex> g(mu); if (!err) err = std::current_exception(); }
    }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < nt; ++t) th.emplace_back(work);
  work();
  for (auto& t : th) t.join();
  if (err) std::rethrow_exception(err);
}
}  // namespace

void run_eigensolver_block(plfem_ctx* ctx, const DevPattern& pat, DevPlan& D, const double* d_vals, const double* d_sigma_node,
                           const BatchDims& bd, std::vector<DesignEig>& des, int ncv, int maxiter, int refine_steps,
                           const double* d_v0, DevBuf<double>& X, EigenResult& res) {
  const int B = bd.nb;
  const int64_t m = bd.moff[B], ld = m;
  const int64_t* moff = bd.d_moff.p;
  cudaStream_t st = ctx->stream;
  int kmax = 0;
  for (const DesignEig& d : des) kmax = std::max(kmax, d.k);
  const int ncvp = std::max(((std::max(ncv, kmax + P) + P - 1) / P) * P, 2 * P);
  for (int b = 0; b < B; ++b)
    if (bd.moff[b + 1] - bd.moff[b] < ncvp + P) throw StatusError(PLFEM_ERR_INVALID, "a design of the batch is smaller than the Lanczos basis");
  const int ldh = ncvp + P, nslots = ncvp / P + 1;
  const int64_t hstride = (int64_t)ldh * P, sstride = (int64_t)ldh * ncvp;
  DevBuf<double> V[2], BV[2], R, U, rt, rdx, opin, h1, h2, Hs, G, Lall, Linv, Sdev;
  DevBuf<int32_t> cstat;
  for (int b = 0; b < 2; ++b) { V[b].alloc(ctx, (size_t)ld * (ncvp + P)); BV[b].alloc(ctx, (size_t)ld * (ncvp + P)); }
  R.alloc(ctx, (size_t)m * P); U.alloc(ctx, (size_t)m * P); rt.alloc(ctx, (size_t)m * P); rdx.alloc(ctx, (size_t)m * P);
  opin.alloc(ctx, (size_t)m * P);
  DevBuf<double> xi;                                    // interleaved solution of the block solve
  xi.alloc(ctx, (size_t)m * P);
  h1.alloc(ctx, (size_t)B * hstride); h2.alloc(ctx, (size_t)B * hstride); Hs.alloc(ctx, (size_t)B * sstride);
  DevBuf<double> hp;                                   // per-slice partial projections
  hp.alloc(ctx, (size_t)B * RSPLIT * hstride);
  G.alloc(ctx, (size_t)B * RSPLIT * P * P); Lall.alloc(ctx, (size_t)B * nslots * P * P); Linv.alloc(ctx, (size_t)B * P * P);
  Sdev.alloc(ctx, (size_t)B * ncvp * ncvp);
  cstat.alloc(ctx, B); cstat.zero();
  const unsigned gm = (unsigned)((m + 255) / 256);
  const dim3 grows((unsigned)((bd.mmax + 255) / 256), B);
  const dim3 gpairs((unsigned)((bd.mmax / 2 + 255) / 256), B);      // kernels whose threads own a pair of rows

  // operator application on P right-hand sides, captured once into a CUDA graph: R = OP(opin)
  // Two versions: the accurate one (rsteps refinement steps) and a RELAXED one with one step fewer.  Inexact-Krylov
  // theory (Bouras/Fraysse, Simoncini/Szyld) allows the error of an operator application to grow like
  // eps / |current Ritz residual|: once every wanted Ritz pair is within `relax_at` of convergence the remaining
  // applications use the relaxed operator (PLFEM_RELAX_AT, 0 = never).
  cudaGraph_t graph = nullptr, graph_lo = nullptr; cudaGraphExec_t gexec = nullptr, gexec_lo = nullptr; int graph_nodes = 0, graph_nodes_lo = 0;
  struct GraphGuard { cudaGraph_t* g; cudaGraphExec_t* e; ~GraphGuard() { if (*e) cudaGraphExecDestroy(*e); if (*g) cudaGraphDestroy(*g); } } guard{&graph, &gexec}, guard_lo{&graph_lo, &gexec_lo};
  DevBuf<uint8_t> refine_node;      // per node: its design takes part in the refinement solve (only set when some designs do not)
  const uint8_t* d_refine = nullptr;
  auto capture_operator = [&](int rsteps, cudaGraph_t& graph, cudaGraphExec_t& gexec, int& graph_nodes) {
    const int before = ctx->launches;
    // designs whose probe found the raw solve accurate skip the (single) refinement solve of the others
    const uint8_t* act = rsteps == 1 ? d_refine : nullptr;
    PLFEM_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    try {
      // interleaved right-hand sides inside the solve: opin -> xi (+ refinement) -> R (columns again)
      run_solve(ctx, D, opin.p, xi.p, P);
      for (int it = 0; it < rsteps; ++it) {
        launch_resid_k(ctx, pat, d_vals, d_sigma_node, xi.p, opin.p, rt.p, P, act);
        run_solve(ctx, D, rt.p, rdx.p, P, act);
        if (it + 1 < rsteps) launch_axpy(ctx, xi.p, rdx.p, m * P);
      }
      deinterleave_add_kernel<<<gm, 256, 0, st>>>(xi.p, rsteps > 0 ? rdx.p : nullptr, ld, m, R.p, act);
      ctx->launches++;
    } catch (...) {
      cudaGraph_t dead = nullptr; cudaStreamEndCapture(st, &dead); if (dead) cudaGraphDestroy(dead);
      throw;
    }
    PLFEM_CUDA(cudaStreamEndCapture(st, &graph));
    graph_nodes = ctx->launches - before; ctx->launches = before;
    PLFEM_CUDA(cudaGraphInstantiate(&gexec, graph, 0));
  };

  // B-orthonormalise the block in R, per design: U = B R, G = R^T U = L L^T, Vn = R L^-T, BVn = U L^-T; L kept in slot `slot`
  auto orthonormalize = [&](double* Vn, double* BVn, int slot) {
    launch_spmm_b(ctx, pat, d_vals, R.p, U.p, P, ld);
    gram_kernel<<<dim3(RSPLIT, B), 256, 0, st>>>(R.p, U.p, ld, moff, G.p);
    chol_kernel<<<B, 32, 0, st>>>(G.p, Lall.p, nslots, slot, Linv.p, cstat.p);
    scale_block_kernel<<<grows, 256, 0, st>>>(R.p, U.p, ld, moff, Linv.p, Vn, BVn);
    ctx->launches += 3;
  };

  // start block: column 0 = v0 (or ones), the others deterministic pseudo-random
  if (d_v0) PLFEM_CUDA(cudaMemcpyAsync(R.p, d_v0, m * sizeof(double), cudaMemcpyDeviceToDevice, st));
  else { fill_kernel<<<gm, 256, 0, st>>>(R.p, m, 1.0); ctx->launches++; }
  start_block_kernel<<<grows, 256, 0, st>>>(R.p, ld, moff);
  ctx->launches++;
  int cur = 0;
  orthonormalize(V[cur].p, BV[cur].p, ncvp / P);     // scratch slot

  // How many refinement steps does the block-LDL^T solve need?  Pivoting is confined to the pivot blocks, so the
  // quality of a raw solve depends on the mesh and the shift: probe it on the start block with one refinement step
  // (rho = |dx| / |x| per design — the size of the first correction, a scale-free estimate of the contraction; the
  // residual norm itself is dominated by the 1e9-size rows of sliver elements) and take the step count that brings
  // rho^(steps+1) below the Lanczos tolerance with a wide margin.  A design whose raw solve is useless (rho >= 0.05)
  // is reported as singular instead of iterating forever.
  int rsteps = refine_steps;
  if (refine_steps < 0) {
    DevBuf<double> nrm;
    nrm.alloc(ctx, (size_t)B * RSPLIT * 2);
    interleave_kernel<<<gm, 256, 0, st>>>(BV[cur].p, ld, m, opin.p);
    run_solve(ctx, D, opin.p, xi.p, P);
    launch_resid_k(ctx, pat, d_vals, d_sigma_node, xi.p, opin.p, rt.p, P);
    run_solve(ctx, D, rt.p, rdx.p, P);                  // the first correction: |dx| / |x| estimates the contraction
    sq_norms_kernel<<<dim3(RSPLIT, B), 256, 0, st>>>(rdx.p, xi.p, moff, nrm.p);
    ctx->launches += 2;
    std::vector<double> hn((size_t)B * RSPLIT * 2);
    nrm.download(hn.data(), hn.size());
    PLFEM_CUDA(stream_wait(st));
    rsteps = 0;
    std::vector<int> need(B, 0);
    for (int b = 0; b < B; ++b) {
      double nr = 0.0, nbv = 0.0;
      for (int sl = 0; sl < RSPLIT; ++sl) { nr += hn[((size_t)b * RSPLIT + sl) * 2]; nbv += hn[((size_t)b * RSPLIT + sl) * 2 + 1]; }
      const double rho = std::sqrt(nr / nbv);
      des[b].solve_residual = rho;
      if (des[b].status != PLFEM_OK) continue;
      if (!(rho < 0.05)) {
        des[b].status = PLFEM_ERR_SINGULAR;
        des[b].err = "the block-LDL^T solve of A - sigma*B is too inaccurate for this mesh and shift (first refinement correction |dx|/|x| = " + std::to_string(rho) +
                     "): a pivot block is numerically singular";
        continue;
      }
      // rho <= 1e-9: the raw solve is already three decades below the Lanczos tolerance (symmetrised pivot-block inverses
      // give 1e-10 on the reference's meshes): no refinement solve at all
      const int mine = rho <= 1e-9 ? 0 : (rho <= 1e-4 ? 1 : (rho <= 2e-3 ? 2 : (rho <= 1e-2 ? 3 : 5)));
      need[b] = mine;
      rsteps = std::max(rsteps, mine);
    }
    // a forest in which the step count is ONE for some designs and zero for others: the refinement solve skips the fronts,
    // rows and corrections of the designs that do not need it (one poorly conditioned mesh doubled the sweeps of
    // the other eleven).  With two or more steps every design takes them all.
    if (rsteps == 1) {
      int nskip = 0;
      for (int b = 0; b < B; ++b) nskip += (need[b] == 0 && des[b].status == PLFEM_OK);
      if (nskip > 0) {
        std::vector<uint8_t> flag((size_t)(m / 2), 1);
        for (int b = 0; b < B; ++b)
          if (need[b] == 0) std::fill(flag.begin() + bd.moff[b] / 2, flag.begin() + bd.moff[b + 1] / 2, (uint8_t)0);
        refine_node.upload(ctx, flag);
        PLFEM_CUDA(stream_wait(st));
        d_refine = refine_node.p;
        res.refine_skipped = nskip;
      }
    }
  }
  res.refine_steps = rsteps;
  capture_operator(rsteps, graph, gexec, graph_nodes);
  static const double relax_at = [] { const char* e = std::getenv("PLFEM_RELAX_AT"); return e ? atof(e) : 0.0; }();
  bool relaxed = false;
  if (relax_at > 0.0 && rsteps >= 1) capture_operator(rsteps - 1, graph_lo, gexec_lo, graph_nodes_lo);

  struct Host {                      // per-design host state of the projected problem
    std::vector<double> Th, T, w, S, Ttail, tail;
    bool expect_final = false;       // the decay of its bounds predicts that the next check finds it converged
    std::vector<int> order;
    bool done = false, newly = false;
    int q_want = 0;
    double worst = 1e300;            // largest relative residual bound among the wanted Ritz pairs at the last check
    double prev_worst = 1e300;       // ... at the check before, `steps_between` block steps earlier
    int best_nconv = -1, stalled = 0; // restarts since the number of converged pairs last grew
  };
  std::vector<Host> hs(B);
  for (Host& h : hs) h.Th.assign((size_t)ncvp * ncvp, 0.0);
  for (int b = 0; b < B; ++b) hs[b].done = des[b].status != PLFEM_OK;     // failed before the iteration: carried along only
  { int nd = 0; for (const Host& h : hs) nd += h.done; if (nd == B) return; }
  std::vector<double> Hh((size_t)B * sstride), Lh((size_t)B * nslots * P * P);
  std::vector<int32_t> cs(B, 0);
  int nb = P, q = 0;                 // basis vectors present; kept Ritz vectors (their block of Th is diagonal)
  int ndone = 0;
  { const int keep = res.refine_steps; res = EigenResult(); res.refine_steps = keep; }
  X.alloc(ctx, (size_t)m * kmax);
  const double eps23 = std::pow(2.220446049250313e-16, 2.0 / 3.0);
  static const int check_every = [] { const char* e = std::getenv("PLFEM_CHECK_EVERY"); return e ? std::max(1, atoi(e)) : 2; }();
  int since_check = 0;
  // A convergence check costs every unfinished design a dense eigensolve on the host (~0.5 ms for a 60 x 60 projection) and
  // the device an idle gap; checking every second block step from 2k basis vectors on was a third of the host time per
  // design.  The worst residual bound of a design falls geometrically between checks, so the next check is scheduled at 60 %
  // of the steps the slowest design is predicted to need (never less than `check_every`, never more than 8).
  int next_gap = check_every, steps_between = 0, n_checks = 0;
  double ms_checks = 0.0;
  auto check_from = [&](int k) { return std::min(ncvp, ((2 * k + P - 1) / P) * P); };
  for (;;) {
    // ---- one block step: image of the last P basis vectors
    const int j0 = nb - P;
    double* Vc = V[cur].p; double* BVc = BV[cur].p;
    interleave_kernel<<<gm, 256, 0, st>>>(BVc + (int64_t)j0 * ld, ld, m, opin.p);
    ctx->launches++;
    PLFEM_CUDA(cudaGraphLaunch(relaxed ? gexec_lo : gexec, st));
    ctx->launches += relaxed ? graph_nodes_lo : graph_nodes;
    res.n_op += P; res.n_block_op++;
    // Two Gram-Schmidt passes.  In exact arithmetic the image of the last block is B-orthogonal to every basis vector except
    // the last two blocks (three-term recurrence) — or, right after a thick restart (j0 == q), the kept Ritz vectors, which
    // couple to the first new block.  The FIRST pass therefore projects against those columns only: it removes the O(1)
    // components (the ones whose removal cancels digits); the SECOND pass runs over the whole basis and removes what
    // rounding left anywhere, now without cancellation — the orthogonality of two full passes for the memory traffic of
    // little more than one (the passes stream the whole basis: 8.4 + 6.4 % of a forest's kernel time for two of them).
    static const bool local_first = [] { const char* e = std::getenv("PLFEM_CGS"); return !(e && e[0] == 'f'); }();    // PLFEM_CGS=full: two full passes
    const int jA = (!local_first || j0 == q) ? 0 : std::max(0, j0 - P);
    const int nA = nb - jA;
    const dim3 gdots((nb + DCOLS - 1) / DCOLS, RSPLIT, B), gsum((nb + 127) / 128, B);
    const dim3 gdotsA((nA + DCOLS - 1) / DCOLS, RSPLIT, B), gsumA((nA + 127) / 128, B);
    dots_block_kernel<<<gdotsA, 256, 0, st>>>(BVc + (int64_t)jA * ld, ld, R.p, moff, nA, hp.p + jA, ldh, hstride);   // CGS pass 1
    sum_slices_kernel<<<gsumA, 128, 0, st>>>(hp.p + jA, ldh, hstride, nA, h1.p + jA);
    update_block_kernel<<<gpairs, 256, 0, st>>>(Vc + (int64_t)jA * ld, ld, h1.p + jA, ldh, hstride, nA, moff, R.p);
    dots_block_kernel<<<gdots, 256, 0, st>>>(BVc, ld, R.p, moff, nb, hp.p, ldh, hstride);          // CGS pass 2
    sum_slices_kernel<<<gsum, 128, 0, st>>>(hp.p, ldh, hstride, nb, h2.p);
    update_block_kernel<<<gpairs, 256, 0, st>>>(Vc, ld, h2.p, ldh, hstride, nb, moff, R.p);
    store_h_kernel<<<gsum, 128, 0, st>>>(h1.p, h2.p, ldh, hstride, nb, jA, Hs.p + (size_t)j0 * ldh, ldh, sstride);
    ctx->launches += 7;
    orthonormalize(Vc + (int64_t)nb * ld, BVc + (int64_t)nb * ld, j0 / P);
    nb += P;
    ++since_check;
    const int c = nb - P;            // basis vectors whose images are known
    const bool full = (c >= ncvp);
    bool due = false;
    for (int b = 0; b < B; ++b) if (!hs[b].done && c >= check_from(des[b].k)) due = true;
    if (!full && !(due && since_check >= next_gap)) continue;
    steps_between = since_check;
    since_check = 0;

    // ---- convergence check on the c x c projected matrix of every unfinished design
    const auto t_check0 = std::chrono::steady_clock::now();
    ++n_checks;
    PLFEM_CUDA(cudaGetLastError());
    Hs.download(Hh.data(), Hh.size());
    Lall.download(Lh.data(), Lh.size());
    cstat.download(cs.data(), B);
    PLFEM_CUDA(stream_wait(st));
    const bool last_chance = full && res.n_restart >= maxiter;
    for_each_design(B, [&](int b) {
      Host& h = hs[b]; DesignEig& de = des[b];
      h.newly = false;
      if (h.done) return;
      if (cs[b]) {
        de.status = PLFEM_ERR_SINGULAR; de.err = "block Lanczos: the residual block lost rank (Cholesky breakdown)";
        h.done = true; return;
      }
      const double* Hb = Hh.data() + (size_t)b * sstride;
      for (int j = q; j < c; ++j)
        for (int i = 0; i <= j; ++i) h.Th[(size_t)j * ncvp + i] = Hb[(size_t)j * ldh + i];
      h.T.assign((size_t)c * c, 0.0);
      for (int j = 0; j < c; ++j)
        for (int i = 0; i <= j; ++i) h.T[(size_t)j * c + i] = h.T[(size_t)i * c + j] = h.Th[(size_t)j * ncvp + i];
      for (double v : h.T)
        if (!std::isfinite(v)) {
          de.status = PLFEM_ERR_SINGULAR; de.err = "Lanczos recurrence produced a non-finite value (shifted operator singular?)";
          h.done = true; return;
        }
      // The residual bounds read only the last P rows of the eigenvector matrix.  A check that is neither a restart nor
      // expected to be the design's last one asks for just those (symmetric_eigen_tail: no accumulation of the
      // tridiagonalising reflectors, rotations on P rows instead of c — about half the cost); if every wanted pair
      // turns out to be converged after all, the full decomposition follows.
      const double* L = Lh.data() + ((size_t)b * nslots + (c - P) / P) * P * P;     // R_last = V_next L^T
      const int k = de.k;
      int nconv = 0;
      double worst = 0.0;
      auto examine = [&](const double* tail, int ldt, int row0) {       // tail(r, col) = Z(c - P + r, col)
        h.order.resize(c);
        std::iota(h.order.begin(), h.order.end(), 0);
        std::stable_sort(h.order.begin(), h.order.end(), [&](int a, int bb) { return std::fabs(h.w[a]) > std::fabs(h.w[bb]); });
        nconv = 0; worst = 0.0;
        for (int i = 0; i < k; ++i) {
          const double* z = tail + (size_t)h.order[i] * ldt + row0;
          double s2 = 0.0;
          for (int a = 0; a < P; ++a) {
            double t = 0.0;
            for (int bb = a; bb < P; ++bb) t += L[bb + a * P] * z[bb];
            s2 += t * t;
          }
          const double bnd = std::sqrt(s2), ref = std::max(eps23, std::fabs(h.w[h.order[i]]));
          if (bnd <= de.tol * ref) nconv++;
          worst = std::max(worst, bnd / ref);
        }
      };
      bool have_vectors = full || last_chance || h.expect_final;
      if (!have_vectors) {
        h.Ttail = h.T;
        symmetric_eigen_tail(c, h.Ttail, h.w, P, h.tail);
        examine(h.tail.data(), P, 0);
        have_vectors = false;
        if (nconv >= k) have_vectors = true;      // converged earlier than predicted: the vectors are needed now
      }
      if (have_vectors) {
        symmetric_eigen(c, h.T, h.w);        // T now holds eigenvectors in columns
        examine(h.T.data(), c, c - P);
      }
      h.prev_worst = h.worst;
      h.worst = worst;
      de.nconv = nconv;
      if (nconv >= k || last_chance) {
        std::vector<int> sel(h.order.begin(), h.order.begin() + k);
        std::sort(sel.begin(), sel.end(), [&](int a, int bb) { return de.sigma + 1.0 / h.w[a] < de.sigma + 1.0 / h.w[bb]; });
        h.S.assign((size_t)c * k, 0.0);
        de.lambda.resize(k); de.theta.resize(k);
        for (int i = 0; i < k; ++i) {
          std::copy(h.T.begin() + (size_t)sel[i] * c, h.T.begin() + (size_t)(sel[i] + 1) * c, h.S.begin() + (size_t)i * c);
          de.theta[i] = h.w[sel[i]];
          de.lambda[i] = de.sigma + 1.0 / h.w[sel[i]];
        }
        de.n_block_op = res.n_block_op; de.n_restart = res.n_restart;
        if (nconv < k) {
          de.status = PLFEM_ERR_NO_CONVERGENCE;
          de.err = "block Lanczos: " + std::to_string(nconv) + " of " + std::to_string(k) + " eigenpairs converged after " + std::to_string(res.n_restart) + " restarts";
        }
        h.done = true; h.newly = true;
      } else if (full && (nconv > h.best_nconv ? (h.best_nconv = nconv, h.stalled = 0) : ++h.stalled) >= 8) {
        // eight restarts without a single newly converged pair: the operator is too inaccurate for the recurrence to make
        // progress (a raw solve the probe judged acceptable but is not) — report it instead of spinning to maxiter
        de.status = PLFEM_ERR_NO_CONVERGENCE;
        de.err = "block Lanczos stagnated: " + std::to_string(nconv) + " of " + std::to_string(k) + " eigenpairs after " +
                 std::to_string(res.n_restart) + " restarts, none newly converged in the last 8 (inaccurate factorisation of A - sigma*B?)";
        de.n_block_op = res.n_block_op; de.n_restart = res.n_restart;
        h.done = true;
      } else if (full) {
        // thick restart: this design wants its q best Ritz vectors (q = ncvp - P*t so that whole blocks fit again)
        const int keep = k + std::min(nconv, (ncvp - k) / 2);
        const int t = std::max(1, (ncvp - keep) / P);
        h.q_want = ncvp - P * t;
      }
    });
    // eigenvectors of the designs that just finished: X_b = V_b S_b
    for (int b = 0; b < B; ++b) {
      if (!hs[b].newly) continue;
      const int k = des[b].k;
      if (c > ROT_MAXIN || hs[b].S.size() > (size_t)ncvp * ncvp) throw StatusError(PLFEM_ERR_INTERNAL, "rotation tile too small");
      double* Sb = Sdev.p + (size_t)b * ncvp * ncvp;      // the design's own slot: no wait between the designs
      PLFEM_CUDA(cudaMemcpyAsync(Sb, hs[b].S.data(), hs[b].S.size() * sizeof(double), cudaMemcpyHostToDevice, st));
      const int64_t mb = bd.moff[b + 1] - bd.moff[b];
      rotate_forest_kernel<<<dim3((unsigned)((mb / 2 + 127) / 128), 1), 128, 0, st>>>(V[cur].p, ld, c, Sb, 0, k, moff + b, X.p, m);
      ctx->launches++;
    }
    ndone = 0;
    for (const Host& h : hs) ndone += h.done;
    ms_checks += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_check0).count();
    if (ndone == B) {
      static const bool timing = std::getenv("PLFEM_TIMING") != nullptr;
      if (timing) fprintf(stderr, "[plfem] block Lanczos: %d block steps, %d convergence checks, %.2f ms of host wall time in them\n", res.n_block_op, n_checks, ms_checks);
      return;
    }
    next_gap = check_every;
    {                                // (a thick restart keeps the wanted Ritz pairs and their residual couplings: the bounds go on
      //                                falling at the measured rate, so the prediction holds across it — a check costs every unfinished
      //                                design a dense eigensolve on the host, 0.3-0.8 ms each, a block step 0.07 ms of device time)
      int need = 0;
      bool known = true;
      for (int b = 0; b < B; ++b) {
        const Host& h = hs[b];
        if (h.done) continue;
        if (!(h.worst < h.prev_worst) || h.prev_worst > 1e299 || steps_between <= 0) { known = false; break; }
        const double rate = std::log(h.worst / h.prev_worst) / steps_between;      // < 0
        need = std::max(need, (int)std::ceil(std::log(des[b].tol / std::max(h.worst, 1e-300)) / rate));
      }
      if (known) next_gap = std::max(check_every, std::min(8, (int)(0.6 * need)));
      for (int b = 0; b < B; ++b) {
        Host& h = hs[b];
        h.expect_final = false;
        if (h.done || !(h.worst < h.prev_worst) || h.prev_worst > 1e299 || steps_between <= 0) continue;
        const double rate = std::log(h.worst / h.prev_worst) / steps_between;
        h.expect_final = std::ceil(std::log(des[b].tol / std::max(h.worst, 1e-300)) / rate) <= next_gap;
      }
    }
    if (!relaxed && gexec_lo) {
      double w = 0.0;
      for (const Host& h : hs) if (!h.done) w = std::max(w, h.worst);
      if (w <= relax_at) { relaxed = true; res.relaxed_from = res.n_block_op; }
    }
    if (!full) continue;
    // ---- thick restart, all designs together: the largest q any unfinished design asks for
    q = 0;
    for (int b = 0; b < B; ++b) if (!hs[b].done) q = std::max(q, hs[b].q_want);
    const int nxt = cur ^ 1;
    {
      const int64_t sst = (int64_t)ncvp * q;
      std::vector<double> S((size_t)B * sst, 0.0);
      for (int b = 0; b < B; ++b) {
        Host& h = hs[b];
        double* Sb = S.data() + (size_t)b * sst;
        if (h.done) {
          for (int i = 0; i < q; ++i) Sb[(size_t)i * ncvp + i] = 1.0;        // finished: carried along, values stay finite
        } else {
          std::fill(h.Th.begin(), h.Th.end(), 0.0);
          for (int i = 0; i < q; ++i) {
            const int col = h.order[i];
            std::copy(h.T.begin() + (size_t)col * ncvp, h.T.begin() + (size_t)(col + 1) * ncvp, Sb + (size_t)i * ncvp);
            h.Th[(size_t)i * ncvp + i] = h.w[col];
          }
        }
      }
      PLFEM_CUDA(cudaMemcpyAsync(Sdev.p, S.data(), S.size() * sizeof(double), cudaMemcpyHostToDevice, st));
      if (ncvp > ROT_MAXIN) throw StatusError(PLFEM_ERR_INVALID, "Lanczos basis larger than the rotation kernel's shared tile");
      const dim3 grot((unsigned)((bd.mmax / 2 + 127) / 128), B);
      rotate_forest_kernel<<<grot, 128, 0, st>>>(V[cur].p, ld, ncvp, Sdev.p, sst, q, moff, V[nxt].p, ld);
      rotate_forest_kernel<<<grot, 128, 0, st>>>(BV[cur].p, ld, ncvp, Sdev.p, sst, q, moff, BV[nxt].p, ld);
      ctx->launches += 2;
      PLFEM_CUDA(stream_wait(st));   // S is a local host buffer
    }
    PLFEM_CUDA(cudaMemcpyAsync(V[nxt].p + (int64_t)q * ld, V[cur].p + (int64_t)ncvp * ld, (size_t)m * P * sizeof(double), cudaMemcpyDeviceToDevice, st));
    PLFEM_CUDA(cudaMemcpyAsync(BV[nxt].p + (int64_t)q * ld, BV[cur].p + (int64_t)ncvp * ld, (size_t)m * P * sizeof(double), cudaMemcpyDeviceToDevice, st));
    cur = nxt; nb = q + P;
    res.n_restart++;
  }
}

// per-mode reductions of ONE design of the forest: rows [row0, row0 + n) of the concatenated pattern / vectors
void run_mode_metrics(plfem_ctx* ctx, const DevPattern& pat, const double* d_vals, const int32_t* d_perm_to_interior,
                      const uint8_t* d_in_core, const double* X, int64_t ldx, int32_t row0, int32_t n,
                      const std::vector<double>& lambda, int k, double* d_out_evecs, double* d_metrics, double* d_resid) {
  cudaStream_t st = ctx->stream;
  const int nchunks = (n + MROWS - 1) / MROWS;
  DevBuf<double> part, lam, scale;
  part.alloc(ctx, (size_t)k * nchunks * NRED);
  lam.upload(ctx, lambda.data(), k);
  scale.alloc(ctx, k);
  mode_partial_kernel<<<dim3(nchunks, (k + MODE_PAIR - 1) / MODE_PAIR), 256, 0, st>>>(row0, row0 + n, pat.rowptr.p, pat.col.p, d_vals, pat.nnz,
                                                                                     d_in_core, (const double2*)X, ldx / 2, lam.p, k, part.p, nchunks);
  mode_final_kernel<<<(k + 63) / 64, 64, 0, st>>>(part.p, nchunks, k, lam.p, d_metrics, d_resid, scale.p);
  ctx->launches += 2;
  if (d_out_evecs) {
    write_evecs_kernel<<<dim3((n + 255) / 256, k), 256, 0, st>>>(row0, n, d_perm_to_interior, (const double2*)X, ldx / 2,
                                                                 scale.p, d_out_evecs);
    ctx->launches++;
  }
  PLFEM_CUDA(cudaGetLastError());
  PLFEM_CUDA(stream_wait(st));   // lambda upload source and the DevBufs above go out of scope
}

}  // namespace plfem
