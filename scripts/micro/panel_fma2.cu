// Micro-benchmark 2 of the sweep tasks' inner loop: a CTA of TW warps multiplies a 128-row panel in shared memory (warp w owns rows
// 128/TW * w ..., R = 4/TW rows per lane) with NR = 4 right-hand sides broadcast from shared memory.  Cycles per panel column for
// the CTA alone on its SM and with several CTAs per SM.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o panel_fma2 panel_fma2.cu
#include <cstdio>
#include <cuda_runtime.h>
constexpr int NR = 4, NCOL = 128;

template <int TW, bool YREG>
__global__ void __launch_bounds__(32 * TW) k(const double* __restrict__ gp, const double* __restrict__ gy, double* out, long long* cyc, int reps) {
  constexpr int R = 4 / TW;
  extern __shared__ double sm[];
  double* panel = sm;                       // per warp: NCOL columns of 32 * R doubles
  double* yv = sm + NCOL * 128;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < NCOL * 128; i += 32 * TW) panel[i] = gp[i];
  for (int i = tid; i < NCOL * NR; i += 32 * TW) yv[i] = gy[i];
  __syncthreads();
  double acc[2][R][NR];
  for (int u = 0; u < 2; ++u) for (int q = 0; q < R; ++q) for (int r = 0; r < NR; ++r) acc[u][q][r] = 0.0;
  const double* base = panel + warp * (NCOL * 32 * R) + lane;
  const long long t0 = clock64();
  for (int rep = 0; rep < reps; ++rep) {
    const double* p = base; const double* y = yv;
    for (int c = 0; c < NCOL; c += 4) {
      double m[4][R]; double2 ya[4], yb[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (YREG) { ya[u] = make_double2(1.0 + u, 2.0); yb[u] = make_double2(3.0, 4.0 + c); }
        else { ya[u] = *reinterpret_cast<const double2*>(y + u * NR); yb[u] = *reinterpret_cast<const double2*>(y + u * NR + 2); }
#pragma unroll
        for (int q = 0; q < R; ++q) m[u][q] = p[32 * q];
        p += 32 * R;
      }
      y += 4 * NR;
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int q = 0; q < R; ++q) {
          acc[u & 1][q][0] = fma(m[u][q], ya[u].x, acc[u & 1][q][0]); acc[u & 1][q][1] = fma(m[u][q], ya[u].y, acc[u & 1][q][1]);
          acc[u & 1][q][2] = fma(m[u][q], yb[u].x, acc[u & 1][q][2]); acc[u & 1][q][3] = fma(m[u][q], yb[u].y, acc[u & 1][q][3]);
        }
    }
  }
  const long long t1 = clock64();
  double s = 0.0;
  for (int u = 0; u < 2; ++u) for (int q = 0; q < R; ++q) for (int r = 0; r < NR; ++r) s += acc[u][q][r];
  out[blockIdx.x * 32 * TW + tid] = s;
  if (tid == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int TW, bool YREG>
void run(const char* name, int ctas_per_sm) {
  cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
  const int grid = pr.multiProcessorCount * ctas_per_sm, reps = 20;
  double *gp, *gy, *out; long long* cyc;
  cudaMalloc(&gp, NCOL * 128 * 8); cudaMalloc(&gy, NCOL * NR * 8); cudaMalloc(&out, (size_t)grid * 32 * TW * 8); cudaMalloc(&cyc, grid * 8);
  cudaMemset(gp, 0, NCOL * 128 * 8); cudaMemset(gy, 0, NCOL * NR * 8);
  const size_t smem = ctas_per_sm == 1 ? (NCOL * 128 + NCOL * NR) * 8 : (NCOL * 128 + NCOL * NR) * 8;
  cudaFuncSetAttribute(k<TW, YREG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k<TW, YREG><<<grid, 32 * TW, smem>>>(gp, gy, out, cyc, reps);
  k<TW, YREG><<<grid, 32 * TW, smem>>>(gp, gy, out, cyc, reps);
  cudaDeviceSynchronize();
  static long long h[8192]; cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < grid; ++i) avg += h[i]; avg /= grid;
  printf("%-44s TW=%d R=%d : %6.1f cycles per 128-row column (%s)\n", name, TW, 4 / TW, avg / (reps * NCOL), cudaGetErrorString(cudaGetLastError()));
  cudaFree(gp); cudaFree(gy); cudaFree(out); cudaFree(cyc);
}

int main() {
  run<1, false>("1 CTA per SM, y broadcast from shared", 1);
  run<2, false>("1 CTA per SM, y broadcast from shared", 1);
  run<4, false>("1 CTA per SM, y broadcast from shared", 1);
  run<1, true>("1 CTA per SM, y in registers (no y loads)", 1);
  run<2, true>("1 CTA per SM, y in registers (no y loads)", 1);
  run<4, true>("1 CTA per SM, y in registers (no y loads)", 1);
  return 0;
}
