// Micro-benchmark 3 of the sweep tasks' inner loop: the same 128-row panel x 4 right-hand sides on the FP64 tensor path
// (mma.sync m8n8k4): warp w owns rows 32 w ..., the panel is stored in fragment order (k-group of 4 columns x row-group of 8 rows =
// 32 consecutive doubles, lane = (row & 7) * 4 + (col & 3)), the right-hand sides are the B fragment (lanes < 16 hold y[k][n], n < 4).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o panel_dmma panel_dmma.cu
#include <cstdio>
#include <cuda_runtime.h>
constexpr int NR = 4, NCOL = 128;

template <int TW>
__global__ void __launch_bounds__(32 * TW) k(const double* __restrict__ gp, const double* __restrict__ gy, double* out, long long* cyc, int reps) {
  constexpr int G = 4 / TW * 1;             // row-groups of 8 per warp = 32 rows / 8 (TW = 4) ... 128 / 8 (TW = 1)
  constexpr int NG = 16 / TW;               // row groups per warp
  extern __shared__ double sm[];
  double* panel = sm;                       // [k-group][row-group (16)][32]
  double* yv = sm + NCOL * 128;             // [col][NR]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < NCOL * 128; i += 32 * TW) panel[i] = gp[i];
  for (int i = tid; i < NCOL * NR; i += 32 * TW) yv[i] = gy[i];
  __syncthreads();
  double c[NG][2];
  for (int g = 0; g < NG; ++g) c[g][0] = c[g][1] = 0.0;
  const double* base = panel + warp * NG * 32 + lane;
  const bool hasb = lane < 16;
  const double* yb = yv + (lane & 3) * NR + (lane >> 2);
  const long long t0 = clock64();
  for (int rep = 0; rep < reps; ++rep) {
#pragma unroll 2
    for (int kg = 0; kg < NCOL / 4; ++kg) {
      const double b = hasb ? yb[kg * 4 * NR] : 0.0;
      double a[NG];
#pragma unroll
      for (int g = 0; g < NG; ++g) a[g] = base[kg * 512 + g * 32];
#pragma unroll
      for (int g = 0; g < NG; ++g)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[g][0]), "+d"(c[g][1]) : "d"(a[g]), "d"(b));
    }
  }
  const long long t1 = clock64();
  double s = 0.0;
  for (int g = 0; g < NG; ++g) s += c[g][0] + c[g][1];
  out[blockIdx.x * 32 * TW + tid] = s;
  if (tid == 0) cyc[blockIdx.x] = t1 - t0;
  (void)G;
}

template <int TW>
void run(int ctas_per_sm) {
  cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
  const int grid = pr.multiProcessorCount * ctas_per_sm, reps = 20;
  double *gp, *gy, *out; long long* cyc;
  cudaMalloc(&gp, NCOL * 128 * 8); cudaMalloc(&gy, NCOL * NR * 8); cudaMalloc(&out, (size_t)grid * 32 * TW * 8); cudaMalloc(&cyc, grid * 8);
  cudaMemset(gp, 0, NCOL * 128 * 8); cudaMemset(gy, 0, NCOL * NR * 8);
  const size_t smem = (NCOL * 128 + NCOL * NR) * 8;
  cudaFuncSetAttribute(k<TW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k<TW><<<grid, 32 * TW, smem>>>(gp, gy, out, cyc, reps);
  k<TW><<<grid, 32 * TW, smem>>>(gp, gy, out, cyc, reps);
  cudaDeviceSynchronize();
  static long long h[8192]; cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < grid; ++i) avg += h[i]; avg /= grid;
  printf("DMMA m8n8k4, fragment-ordered panel  TW=%d (%2d row groups per warp): %6.1f cycles per 128-row column (%s)\n", TW, 16 / TW,
         avg / (reps * NCOL), cudaGetErrorString(cudaGetLastError()));
  cudaFree(gp); cudaFree(gy); cudaFree(out); cudaFree(cyc);
}

int main() {
  run<1>(1); run<2>(1); run<4>(1);
  return 0;
}
